cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp9b.log
: > $L
timeout 600 python -m pytest tests/test_gpu_packed.py -m gpu -x -q -k "host_buffer" 2>&1 | tail -2 >> $L
P='import json,sys; d=json.loads(sys.stdin.read()); e=d["e2e"]; print(e["value"], e["ms_per_step"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"])'
for ch in 8 8 12 16; do
  echo -n "e2e chunk $ch (short first chunk): " >> $L
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --e2e-steps 6 --e2e-chunk $ch 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
cat $L
