# r02 job 19 (8 GPUs): CUPTI timeline of two resident steps of the 8-image run on rank 0 (what the 3.8 ms over the 1-image step are)
set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scripts/e2e_trace.py --mode step --min-us 40 --out gpurun_out/r02t_step_trace_n8.json > gpurun_out/r02t_step_trace_n8.txt 2> gpurun_out/r02t_step_trace_n8.err; echo "rc=$?"
tail -3 gpurun_out/r02t_step_trace_n8.err
gzip -f gpurun_out/r02t_step_trace_n8.json; ls -la gpurun_out/r02t_step_trace_n8.json.gz
