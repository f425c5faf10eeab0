mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json | cut -c1-600; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 600 python scripts/bench_c3_c4.py > gpurun_out/bench_c3_c4.json 2> gpurun_out/bench_c3_c4.err; echo "c3c4 exit $?"; cat gpurun_out/bench_c3_c4.json; tail -5 gpurun_out/bench_c3_c4.err
