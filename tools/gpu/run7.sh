mkdir -p gpurun_out
( time python bench.py > gpurun_out/r02a_bench_c3_full.json 2> gpurun_out/r02a_bench.err ) 2> gpurun_out/r02a_time.txt
tail -3 gpurun_out/r02a_time.txt; grep -v "^\*\|OMP" gpurun_out/r02a_bench.err | tail -5
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_reference.json 2> gpurun_out/r02a_reference.err ) 2>> gpurun_out/r02a_time.txt
tail -3 gpurun_out/r02a_time.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/san_case.py 0.1 > gpurun_out/r02a_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02a_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/san_case.py 0.05 > gpurun_out/r02a_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02a_racecheck.log
