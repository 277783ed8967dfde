# end-of-round check on one GPU: all GPU tests, smoke(), the bench line, the ncu launch list of the bench command
out=gpurun_out/${1:-r2n}
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; tail -3 $out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
python bench.py > $out/bench.json 2> $out/bench.err; tail -c 300 $out/bench.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 400 $out/bench_reference.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 1 --warmup 1 > $out/ncu_bench.log 2>&1; tail -2 $out/ncu_bench.log | cut -c1-200
