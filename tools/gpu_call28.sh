mkdir -p gpurun_out
EML_AB_SET=planes timeout 600 python tools/layer_ab.py > gpurun_out/layer_ab_planes.log 2>&1; head -20 gpurun_out/layer_ab_planes.log; tail -1 gpurun_out/layer_ab_planes.log
