for v in vA; do
  cp w2dynamics_b200/libw2b_cthyb.so /tmp/base.so
  cp w2dynamics_b200/libw2b_$v.so w2dynamics_b200/libw2b_cthyb.so
  touch w2dynamics_b200/libw2b_cthyb.so
  python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['kernel_ms'], d['config']['error_flags'])"
  tail -2 gpurun_out/err.txt
  cp /tmp/base.so w2dynamics_b200/libw2b_cthyb.so
done
