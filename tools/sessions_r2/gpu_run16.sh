cd $GRAFT_REPO_ROOT
# final state check: smoke, the whole GPU suite, the default bench line and the reference arm, exactly as the driver runs them
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
( time python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real
python tools/show_bench.py gpurun_out/bench_final.json | cut -c1-420
( time python bench.py --impl reference > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err ) 2>&1 | grep real
python tools/show_bench.py gpurun_out/bench_final_reference.json | cut -c1-200
