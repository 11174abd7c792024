set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r29_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r29_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r29_ref.json 2> gpurun_out/r29_ref.err; echo "ref rc=$?"
python bench.py > gpurun_out/r29_bench.json 2> gpurun_out/r29_bench.err; echo "bench rc=$?"
cat gpurun_out/r29_bench.json
