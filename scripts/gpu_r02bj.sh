#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/r02bj_sanitizer_memcheck.log python scripts/sanitize_argmin.py > gpurun_out/r02bj_sanitizer_stdout.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r02bj_sanitizer_stdout.log; tail -3 gpurun_out/r02bj_sanitizer_memcheck.log
} 2>&1 | tee gpurun_out/r02bj.log
