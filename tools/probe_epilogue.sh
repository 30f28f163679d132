for cfg in "1000000 4096" "10000000 1024" "10000000 256"; do
  set -- $cfg
  for mode in 1 2; do
    for skip in 0 2 1; do
      echo "== rows=$1 nq=$2 mode=$mode skip=$skip"
      IVR_MMA_MODE=$mode IVR_MMA_DEBUG_SKIP_EPILOGUE=$skip timeout 120 python tools/search_one.py $1 512 $2 100 2 2 4 2>&1 | tail -1
    done
  done
done
