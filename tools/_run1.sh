python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2aa_tests.log
python bench.py > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2aa_mt_launches.csv python tests/bench_mt.py --config 3 --reps 1 > gpurun_out/r2aa_ncu_mt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mt_terms2 -s 10 -c 1 -o gpurun_out/r2aa_mtterms2 -f python tests/bench_mt.py --config 3 --reps 1 > gpurun_out/r2aa_ncu_full.log 2>&1
cat gpurun_out/r2aa_tests.log; cut -c1-300 gpurun_out/r2aa_bench.json; tail -3 gpurun_out/r2aa_bench.err
