#!/bin/bash
# ncu --set full captures of the KW=2 (config 3) and KW=4 (config 4) kernel pairs and of the K0r microbenchmark
set -u
mkdir -p gpurun_out
{
  export TSXC_REGION_LOG2=25
  for w in c3 c4; do
    if [ $w = c3 ]; then export TSXC_CHUNK_KEYS=800000000; else export TSXC_CHUNK_KEYS=400000000; fi
    CMD="python bench.py --workload $w --scale 0.25 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
    timeout 300 $CMD > gpurun_out/kw_${w}_plain.log 2>&1; echo "$w plain rc=$?"; grep "timed steps" gpurun_out/kw_${w}_plain.log
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_part_reads|k_insert_keys" -c 2 -o gpurun_out/r02_${w}_quarter -f $CMD > gpurun_out/kw_${w}_ncu.log 2>&1
    echo "$w ncu rc=$?"; tail -2 gpurun_out/kw_${w}_ncu.log
  done
  unset TSXC_REGION_LOG2 TSXC_CHUNK_KEYS
  cat > /tmp/k0r_one.py <<'PY'
import ctypes as C, sys
sys.path.insert(0, ".")
import tsxcount_b200 as tsx
lib = tsx._lib.load()
hm = tsx.TSXHashMapCUDA(32, 0, 31)
tb = hm.stats()["table_bytes"]; rb = 128 << 20
ops = int(rb // 32 * 0.93)
for mode in (2, 5):
    ms = C.c_float(0)
    tsx._lib.check(lib.tsxc_k0_region_sweep(hm.handle, tb, rb, ops, 1024, mode | (6 << 8), C.byref(ms)), hm.handle)
    print(mode, ops * (tb // rb) / ms.value / 1e6, "G ops/s")
hm.close()
PY
  timeout 600 ncu --set full --clock-control none -k regex:"k_k0_region_sweep" -c 2 -o gpurun_out/r02_k0r_128MiB -f python /tmp/k0r_one.py > gpurun_out/kw_k0r_ncu.log 2>&1
  echo "k0r ncu rc=$?"; tail -3 gpurun_out/kw_k0r_ncu.log
} 2>&1 | tee gpurun_out/s2_ncu_kw.txt
