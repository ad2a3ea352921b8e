# usage: bash tools/ab_sweep.sh "kernel pair stages depth" ...   (A/B sweep of the fused DCN forward knobs)
for cfg in "$@"; do
  set -- $cfg
  echo "== kernel=$1 pair=$2 stages=$3 depth=$4"
  KGDET_UMMA_KERNEL=$1 KGDET_UMMA_PAIR=$2 KGDET_UMMA_STAGES=$3 KGDET_UMMA_DEPTH=$4 timeout 120 python tools/dcn_ab.py 2>&1 | tail -2
done
