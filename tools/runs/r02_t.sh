# round 2, GPU call T: compute-sanitizer memcheck over the whole -m gpu suite and smoke() on the final build
mkdir -p gpurun_out
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/t_memcheck_tests.log 2>&1
echo "exit $?" >> gpurun_out/t_memcheck_tests.log
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python __graft_entry__.py smoke ) > gpurun_out/t_memcheck_smoke.log 2>&1
echo "exit $?" >> gpurun_out/t_memcheck_smoke.log
tail -5 gpurun_out/t_memcheck_tests.log; tail -5 gpurun_out/t_memcheck_smoke.log
