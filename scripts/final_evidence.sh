#!/bin/bash
# Round-end evidence on the GPU box, one call: tests, parity report, bench lines, ncu launch list + full captures, sanitizer.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_pytest.log; cat gpurun_out/final_pytest.log
timeout 600 python tests/tools/parity_report.py --out gpurun_out/parity_report.txt > /dev/null 2>&1; tail -3 gpurun_out/parity_report.txt
timeout 700 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_final.json; cut -c1-400 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json; cut -c1-300 gpurun_out/bench_ref.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --ffis 128 --no-cpu --no-prepare --no-configs --e2e-ffis 32 > gpurun_out/bench_under_ncu.log 2>&1
timeout 700 scripts/prof_fit.sh r2 16 33 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_bkgshe" -f -o /tmp/prof_she python scripts/dev_she.py 8 > gpurun_out/prof_she.log 2>&1
ncu -i /tmp/prof_she.ncu-rep --page raw --csv > gpurun_out/prof_she_raw.csv
for t in memcheck racecheck; do timeout 300 compute-sanitizer --tool $t python scripts/sanitize_run.py 2>&1 | grep -E "SUMMARY" ; done > gpurun_out/sanitizer.log; cat gpurun_out/sanitizer.log
cuobjdump -sass photometry_b200/lib/libtbk.so | grep -E "UBLKCP|UTMALDG|UTMASTG|SYNCS|LDGSTS" | sort | uniq -c > gpurun_out/sass_grep.txt; cat gpurun_out/sass_grep.txt
ls -la gpurun_out | wc -l; du -sh gpurun_out
