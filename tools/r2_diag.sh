set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,mig.mode.current,compute_mode --format=csv
python - <<'PY' > gpurun_out/r2d_diag.log 2>&1
import importlib, torch, ctypes
print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0))
pkg = importlib.import_module("low-cost-mocap_b200")
import numpy as np
for C in (4, 8, 16):
    try:
        ctx = pkg.MocapContext(C, max_roots=64 if C == 8 else 16)
        print(C, "created")
        K = np.array([[600.0,0,320],[0,600,240],[0,0,1]])
        poses, _ = pkg.synth.make_rig(C)
        ctx.set_cameras([K]*C, poses)
        fr = torch.zeros((2, C, 480, 640), dtype=torch.uint8, device="cuda")
        out = ctx.pipeline(fr); torch.cuda.synchronize(); print(C, "pipeline ok")
    except Exception as e:
        print(C, "FAILED", e)
PY
cat gpurun_out/r2d_diag.log
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "ba_device or config3 or golden" 2>&1 | tail -8
