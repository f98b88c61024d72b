"""include/pqp_multi.h: the single-process multi-GPU path behind the C ABI.

CPU: symbols exported, creation fails loudly without a device. GPU: the plain-C test program
(tests/cpp/multi_test.c) shards a batch over 1 device (any GPU box) and over 2 devices (when the box
has them: `gpurun --gpus 2`) and compares with the single-device call bit for bit, gathered table on
every device included."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from path_optimizer_2_b200 import abi, solver, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "multi_test")


def _build():
    solver.build_library()
    src = os.path.join(ROOT, "tests", "cpp", "multi_test.c")
    deps = [src, os.path.join(ROOT, "include", "pqp_multi.h"), solver.LIB_PATH]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-I/usr/local/cuda/include", "-o", BIN, src, solver.LIB_PATH,
                               "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.dirname(solver.LIB_PATH),
                               "-Wl,-rpath,/usr/local/cuda/lib64"])
    return BIN


def _write(path, hb):
    with open(path, "wb") as f:
        np.array([hb.batch, hb.n_max], dtype=np.int32).tofile(f)
        hb.knots.tofile(f)
        hb.inst.tofile(f)
        hb.n.tofile(f)


def test_multi_symbols_and_no_device_error():
    from path_optimizer_2_b200 import multi
    L = multi._lib()
    assert all(hasattr(L, s) for s in multi.EXPORTED_SYMBOLS)
    declared = set()
    import re
    with open(os.path.join(ROOT, "include", "pqp_multi.h")) as f:
        for m in re.finditer(r"^\s*(?:int|const char \*)\s*(pqp_\w+)\s*\(", f.read(), re.M):
            declared.add(m.group(1))
    assert declared == set(multi.EXPORTED_SYMBOLS)
    _build()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(solver.PqpError) as e:
        multi.MultiGpuSolver(n_max=60, batch_max=8, n_devices=2)
    assert e.value.code == abi.PQP_E_NO_DEVICE


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [1, 2])
def test_multi_device_batch_through_the_c_abi(tmp_path, devices):
    import torch
    if torch.cuda.device_count() < devices:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (devices, devices))
    exe = _build()
    hb = synthetic.make_batch(3, 1001, 120, ragged=True)  # odd size: the last shard is shorter
    p = str(tmp_path / "batch.bin")
    _write(p, hb)
    out = subprocess.run([exe, p, str(devices)], capture_output=True, text=True, check=True).stdout.strip()
    out = out.splitlines()[-1]  # NCCL may print its version banner first
    assert out.startswith("ok devices %d batch 1001" % devices), out
    print(out)


@pytest.mark.gpu
def test_pack_results_kernel():
    import torch
    from path_optimizer_2_b200 import multi
    hb = synthetic.make_batch(3, 300, 60)
    sv = solver.PathQpSolver(abi.default_params(), n_max=60, batch_max=300)
    r = sv.solve(hb)
    dev = torch.device("cuda", 0)
    d_cost, d_status, d_iters = (torch.from_numpy(v).to(dev) for v in (r.cost, r.status, r.iters))
    packed = torch.zeros(300 * 16, dtype=torch.uint8, device=dev)
    multi.pack_results_device(sv, 300, d_cost.data_ptr(), d_status.data_ptr(), d_iters.data_ptr(), packed.data_ptr(),
                              stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    rec = packed.cpu().numpy().view(multi.RESULT_REC)
    assert np.array_equal(rec["cost"], r.cost) and np.array_equal(rec["status"], r.status) and np.array_equal(rec["iters"], r.iters)
    sv.close()
