for n in 16384 8192 4096; do
  for t in 0 1; do
    echo "n=$n tracked=$t: $(MCHRG_B200_POTRF_TRACKED=$t python tools/prof_potrf.py $n 3 2>&1 | tail -1)"
  done
done
for g in 256 1024; do echo "n=16384 tracked group=$g: $(MCHRG_B200_POTRF_TRACKED=1 MCHRG_B200_DIST_GROUP=$g python tools/prof_potrf.py 16384 3 2>&1 | tail -1)"; done
