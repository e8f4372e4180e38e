run() { echo "[$*] "; env "$@" python tools/kernel_times.py 60 40 2>&1 | grep -E "cfg=|force|density|integrate|reorder|hash|scan|insert|clear|wave" | sed "s/  */ /g"; }
run CWA_BENCH_GRID=h_y30 CWA_NB_CONFIG=10
run CWA_BENCH_GRID=h_y30 CWA_NB_CONFIG=1
run CWA_NB_CONFIG=10
