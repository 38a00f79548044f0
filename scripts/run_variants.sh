for thr in 512 640 768; do
  echo "== threads $thr"
  PPGPU_K2P_THREADS=$thr python scripts/fam_times.py synthetic_30_6_40_s0 5 2 2>&1 | grep -E "^rep 1|k2a_relax" | tail -2 | cut -c1-200
done
