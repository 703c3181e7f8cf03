timeout 500 python tools/coop_check.py golden c5 2>&1 | grep -E "bad [1-9]|c5 tile|c2_static_rand |Error|error"
SPHRT_WALKER=coop python tools/kbench.py c5:8 > gpurun_out/kb.log 2>&1 && SPHRT_WALKER=coop ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,launch__block_size,launch__registers_per_thread --clock-control none -k regex:walk_coop -c 2 --csv --log-file gpurun_out/l.csv python tools/kbench.py c5:2 > /dev/null 2>&1
cat gpurun_out/kb.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/l.csv')) if len(r)>10]
h=rows[0]; k=h.index('Kernel Name'); m=h.index('Metric Name'); v=h.index('Metric Value')
for r in rows[1:]:
    print(('fwd' if 'Forward' in r[k] else 'adj'), r[m], r[v])
PY
