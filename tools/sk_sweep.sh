# cluster split-K dispatch sweep: in-graph time of RMVPE and HuBERT alone (tools/stage_timing.py ONLY_RMVPE=1)
export ONLY_RMVPE=1
for cfg in "4 16 0" "8 16 0" "8 8 0" "4 8 0" "8 8 1" "4 16 1" "8 4 0" "2 16 0"; do set -- $cfg
  echo "SK_MAX=$1 SK_MINKB=$2 SK_BN64=$3"; RVCB_SK_MAX=$1 RVCB_SK_MINKB=$2 RVCB_SK_BN64=$3 timeout 200 python tools/stage_timing.py 2>&1 | grep "RMVPE\|HuBERT"
done
