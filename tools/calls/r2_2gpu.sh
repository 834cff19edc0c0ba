# two-GPU call: NCCL ingest / egress equals the single-GPU result, then the bench at N = 2 (both arms' launch lines as the driver uses them)
O=gpurun_out/r2_2gpu; mkdir -p $O; cd /root/repo
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -rs > $O/tests_multi_gpu.log 2>&1; echo "rc=$?" >> $O/tests_multi_gpu.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
tail -n 3 $O/tests_multi_gpu.log; tail -c 1500 $O/bench_2gpu.json; tail -n 5 $O/bench_2gpu.err
