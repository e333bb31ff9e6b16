# ncu full capture of selected kernels: KREGEX, SKIP, COUNT, OUT
B="python bench.py --steps 1 --warmup 1 --inflight 1 --repeats 1 --no-cpu-baseline --no-e2e --no-graph"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-fuse_tc}" -s ${SKIP:-12} -c ${COUNT:-1} -o gpurun_out/${OUT:-prof_one} -f $B > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log; ls -la gpurun_out/${OUT:-prof_one}.ncu-rep
