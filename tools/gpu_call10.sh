#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_call10_bench.log
run() { # name lib spt
  for cfg in "--config c2 --arith 0" "--config c2 --arith 1" "--config c5 --arith 1"; do
    echo -n "$1 spt=$3: " >> gpurun_out/r2_call10_bench.log
    PTB_SPT=$3 PTB_LIB=$2 python tools/quick_bench.py $cfg --batch 8 --launches 3 2>&1 | grep "Msegments" | cut -c1-90 >> gpurun_out/r2_call10_bench.log
  done
}
for spt in 8 4 2; do run default szakdolgozat_pathtracer_b200/libptb.so $spt; run cg build/var_cg/libptb.so $spt; done
cat gpurun_out/r2_call10_bench.log
# sanitizer probe (memcheck + racecheck on a small launch of every pipeline)
cat > /tmp/san.py <<'PY'
import sys; sys.path[:0]=['.','tests','tools']
import numpy as np, make_assets, szakdolgozat_pathtracer_b200 as ptb
from scenes import load_config
ctx=ptb.Context(0); sc=load_config(ptb, make_assets, "c2"); h,_=ctx.accel_build(sc)
W,H=96,64; n=W*H; a,f=ctx.alloc(n*16),ctx.alloc(n*4); ctx.memset(a,0,n*16)
for pipe in (1,2,3,4):
    for ar in (0,1):
        if ar and pipe not in (2,3): continue
        p=ptb.make_params(W,H,subframe_index=0,dof=True,eye=(0.0,1.2,3.2),lookat=(0.0,0.7,0.0)); p.accum_buffer,p.frame_buffer,p.handle=a,f,h
        ctx.launch(p, ptb.default_render_cfg(spp_per_launch=2,max_depth=4,subframes_per_launch=2,pipeline=pipe,arith_mode=ar)); ctx.synchronize()
print("sanitizer workload done", ctx.launch_stats().segments)
PY
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/r2_sanitizer_$tool.txt | cut -c1-200
done
