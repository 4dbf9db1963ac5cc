for c in 2 3; do
  SE3MPC_CTAS_PER_SM=$c SE3MPC_LIB=$PWD/build/tune/lib_q_h2b.so python profiles/ab_one.py "q_h2b c$c" 65536 2>&1 | head -1
done
