#!/bin/bash
# BASELINE.json config 4: 3840x2160 4xAA demo scene, tile rows sharded over 1/2/4/8 GPUs of one
# context (NVLink framebuffer gather), through the engine API; reference on all host cores beside it.
N=$(nproc)
A="-s demo03 -x 3840 -y 2160 -a 2 -g -d 16 -f 60 -w 5 -t $N"
echo "reference ${N} threads: $(oracle/_ref/qr_ref_harness $A | python -c 'import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j["ms_med"])') ms"
for d in 0 0,1 0,1,2,3 0,1,2,3,4,5,6,7; do
  s=$(QR_B200_DEVICES=$d build/qr_b200_harness $A | python -c 'import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j["ms_med"])')
  p=$(QR_B200_DEVICES=$d QR_B200_PIPELINE=1 build/qr_b200_harness $A | python -c 'import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j["ms_med"])')
  # engine tiling off, tile lists built on the device
  t=$(QR_B200_DEVICES=$d QR_B200_PIPELINE=1 QR_B200_EXPECT_DEVICE_TILING=1 build/qr_b200_harness $A -p 0x0230FFB9 | python -c 'import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j["ms_med"])')
  echo "GPUs $d: synchronous $s ms, pipelined $p ms, pipelined + device-side tiling $t ms"
done
