O=gpurun_out/r2ab; mkdir -p $O; cd /root/repo
timeout 150 python -m pytest tests/test_zz_model_route.py tests/test_experiment_dir.py -m "gpu or not gpu" -q -k "c99" > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
tail -n 30 $O/tests.log | cut -c1-400
