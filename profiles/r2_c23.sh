# 2 vs 3 resident CTAs per SM (255 vs 168 registers) by launch size: SE3MPC_CTAS_PER_SM forces the instance
for big in 9216 16384 32768; do for c in 2 3; do
  SE3MPC_CTAS_PER_SM=$c python profiles/ab_one.py "c$c" $big 2>&1 | head -1 | sed 's/.*|//' | sed "s/^/c$c /"
done; done
