# what the driver runs at round end, in one go (on a GPU box)
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; head -c 400 gpurun_out/final_ref.json; echo
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err; head -c 300 gpurun_out/final_bench.json; echo
