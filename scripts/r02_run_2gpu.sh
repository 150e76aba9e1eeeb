# round 2, 2-GPU call: multi-GPU tests on hardware (torchrun + NCCL), bench --gpus 2
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q --no-header -rf --timeout 900 -k "multi or two_gpus or several_gpus or second_device or staged_table" ) > gpurun_out/r02_tests_2gpu.log 2>&1
tail -12 gpurun_out/r02_tests_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -c 2500 gpurun_out/r02_bench_2gpu.json; tail -3 gpurun_out/r02_bench_2gpu.err
