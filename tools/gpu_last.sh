mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_densenet_gpu.py -q --timeout 200 -p no:cacheprovider -k "channel_plane or golden or independence" 2>&1 | tail -3
timeout 120 python tools/fwd_time.py 256 2>&1 | tail -1
