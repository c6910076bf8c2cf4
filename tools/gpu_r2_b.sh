#!/bin/bash
mkdir -p gpurun_out
timeout 300 build/umma_probe 50000 200 2048 > gpurun_out/umma_probe.txt 2>&1; echo "probe exit $?" >> gpurun_out/umma_probe.txt
timeout 600 python -m pytest tests/test_zz_bcpd.py -m gpu -q --runxfail -x > gpurun_out/bcpd_tests.txt 2>&1; echo "exit $?" >> gpurun_out/bcpd_tests.txt
timeout 900 python tools/diag_r2.py > gpurun_out/diag_r2.txt 2>&1; echo "exit $?" >> gpurun_out/diag_r2.txt
cat gpurun_out/umma_probe.txt; tail -5 gpurun_out/bcpd_tests.txt; cat gpurun_out/diag_r2.txt
