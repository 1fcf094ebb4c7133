set -e
run() { HB2_FAST_FLAGS="$1" python -m hamers_b200.build --force > /dev/null; echo "== flags: $1"; QUICK_NCU=0 bash tools/quick.sh 512; }
run ""
run "-DHB2_STREAM_HINTS=1"
run "-DHB2_PREFETCH_R=0"
run "-DHB2_STREAM_HINTS=1 -DHB2_PREFETCH_R=0"
