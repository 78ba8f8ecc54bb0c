set -x
python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; tail -2 gpurun_out/f_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; tail -c 300 gpurun_out/f_ref.json
python bench.py --steps 20 --warmup 5 > gpurun_out/f_n1.json 2> gpurun_out/f_n1.err; tail -c 200 gpurun_out/f_n1.err
python bench.py --steps 20 --warmup 5 --e2e-stream cblocks --e2e-output dense --no-cpu-baseline > gpurun_out/f_n1_oldformats.json 2> /dev/null
python bench.py --steps 10 --warmup 3 --coverage 100 --bin-width 50 --no-cpu-baseline > gpurun_out/f_config4.json 2> gpurun_out/f_config4.err; tail -c 200 gpurun_out/f_config4.err
python bench.py --workload cohort --samples 64 --samples-per-call 32 --steps 2 --warmup 1 > gpurun_out/f_cohort64.json 2> gpurun_out/f_cohort64.err; tail -c 200 gpurun_out/f_cohort64.err
python tools/bench_packer.py --reads 8000000 > gpurun_out/f_packer.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_expand_wire|k_tally|k_bin_tiles2' -c 12 -o gpurun_out/f_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-groups 4 > gpurun_out/f_ncu2.log 2>&1
ls -la gpurun_out/f_full.ncu-rep
