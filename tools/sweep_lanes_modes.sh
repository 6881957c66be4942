for cfg in "12 8" "8 8" "16 8" "12 4" "12 12" "20 8"; do set -- $cfg; echo "refill=$1 wait=$2"; CRAY_REFILL_LANES=$1 CRAY_WAIT_LANES=$2 MODES="${MODES:-f32}" bash tools/bench_modes.sh; done
