run() { echo -n "$1 => "; env $1 python bench.py --steps 50 --skip-cpu --skip-bm25 --skip-e2e --skip-extra 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; }
run "SDBG_GROUPBY_TMA_STAGES=4"
run "SDBG_GROUPBY_TMA_STAGES=3"
run "SDBG_GROUPBY_TMA_STAGES=2"
run "SDBG_GROUPBY_TMA_STAGES=2 SDBG_GROUPBY_TMA_CTAS=3"
run "SDBG_GROUPBY_TMA_STAGES=2 SDBG_GROUPBY_TMA_CTAS=4"
run "SDBG_GROUPBY_TMA_STAGES=3 SDBG_GROUPBY_TMA_CTAS=2"
run "SDBG_GROUPBY_TMA_STAGES=4"
run "SDBG_GROUPBY_TMA_STAGES=2 SDBG_GROUPBY_TMA_CTAS=3"
