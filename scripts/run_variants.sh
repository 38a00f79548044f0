for cfg in "16 512" "16 640" "16 768" "16 1024" "32 768" "32 896" "32 1024"; do
  set -- $cfg
  echo "== lanes $1 threads $2"
  PPGPU_K2P_LANES=$1 PPGPU_K2P_THREADS=$2 python scripts/fam_times.py synthetic_30_6_40_s0 5 2 2>&1 | grep -E "^rep 1|k2a_relax" | tail -2 | cut -c1-260
done
