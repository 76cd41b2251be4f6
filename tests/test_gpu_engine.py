"""Fabric PS engine on one GPU (ps + worker colocated) vs a pure-PyTorch oracle of the same protocol."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle_step(p, x, y, opt, state, t, precision="bf16"):
    """fp32 PyTorch reference of one sync step.  precision="bf16": GEMM operands rounded to bf16 (what the bf16 kernels
    compute); "fp32": NO rounding anywhere -- the reference model itself (/root/reference/distributed_mnist.py:98-113),
    which the tf32 engine is held against."""
    r = (lambda v: v.bfloat16().float()) if precision == "bf16" else (lambda v: v)
    h = torch.relu(r(x) @ r(p["hid_w"]) + p["hid_b"])
    h16 = r(h)
    logits = h16 @ r(p["sm_w"]) + p["sm_b"]
    prob = torch.softmax(logits, -1)
    loss = -(y * torch.log(torch.clamp(prob, 1e-10, 1.0))).sum()
    dl = prob - y
    g = {"sm_w": h16.t() @ dl, "sm_b": dl.sum(0)}
    dh = r((dl @ r(p["sm_w"]).t()) * (h16 > 0))
    g["hid_b"] = ((dl @ r(p["sm_w"]).t()) * (h16 > 0)).sum(0)
    g["hid_w"] = r(x).t() @ dh
    for k in p:
        if opt["kind"] == "sgd":
            p[k] = p[k] - opt["lr"] * g[k]
        else:
            m, v = state.setdefault(k, (torch.zeros_like(p[k]), torch.zeros_like(p[k])))
            lr_t = opt["lr"] * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
            m = 0.9 * m + 0.1 * g[k]
            v = 0.999 * v + 0.001 * g[k] * g[k]
            p[k] = p[k] - lr_t * m / (v.sqrt() + 1e-8)
            state[k] = (m, v)
    return float(loss)


@pytest.mark.parametrize("kind", ["sgd", "adam", "momentum"])
def test_colocated_tf32_engine_matches_pure_fp32_oracle(kind):
    """Default precision (fp32 storage, TF32 MMAs, one-kernel worker step) against the UNROUNDED fp32 model: 100 steps, loss
    trajectory within 1e-3 relative, parameters within TF32 noise."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.ops import cuda_lib
    from distributed_tensorflow_b200.parallel.fabric import Fabric
    from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    torch.cuda.set_device(0)
    opt = {"kind": kind, "lr": {"sgd": 0.0005, "momentum": 0.0002, "adam": 0.001}[kind], "momentum": 0.9}
    eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer=opt, seed=3), Fabric(1, {0: 0}))
    assert eng.tf32 and eng.launches_per_worker_step() == 1
    eng.init_params()
    p = {k: v.clone().double() for k, v in eng.state_dict().items() if k in ("hid_w", "hid_b", "sm_w", "sm_b")}
    xs, ys = synthetic_mnist(10000, seed=5)
    n0 = cuda_lib.launch_count()
    state, losses, ref_losses = {}, [], []
    steps = 100
    for t in range(1, steps + 1):
        x = torch.from_numpy(xs[(t - 1) * 100:t * 100])
        y = torch.from_numpy(ys[(t - 1) * 100:t * 100])
        losses.append(eng.step(x.pin_memory(), y.pin_memory()))
        if kind == "momentum":
            # TF momentum: accum = m * accum + g ; var -= lr * accum
            xd, yd = x.double(), y.double()
            h = torch.relu(xd @ p["hid_w"] + p["hid_b"])
            prob = torch.softmax(h @ p["sm_w"] + p["sm_b"], -1)
            ref_losses.append(float(-(yd * torch.log(torch.clamp(prob, 1e-10, 1.0))).sum()))
            dl = prob - yd
            dh = (dl @ p["sm_w"].t()) * (h > 0)
            g = {"sm_w": h.t() @ dl, "sm_b": dl.sum(0), "hid_b": dh.sum(0), "hid_w": xd.t() @ dh}
            for k in p:
                acc = state.get(k, torch.zeros_like(p[k])) * 0.9 + g[k]
                state[k] = acc
                p[k] = p[k] - opt["lr"] * acc
        else:
            ref_losses.append(_oracle_step(p, x.double(), y.double(), opt, state, t, precision="fp32"))
    eng.check_errors()
    assert cuda_lib.launch_count() - n0 == 2 * steps                 # ONE worker kernel + ONE ps kernel per step
    sd = eng.state_dict()
    assert int(sd["global_step"]) == steps
    # measured on a B200: max relative deviation of the per-step batch-sum loss from the unrounded fp32 model over 100 steps
    # = 2.4e-3 (TF32 keeps 10 mantissa bits of x and W1; the bf16 path sits at ~2e-2 on the same trajectory)
    np.testing.assert_allclose(losses, ref_losses, rtol=5e-3, atol=0 if kind == "sgd" else 1e-2)
    for k in ("hid_w", "hid_b", "sm_w", "sm_b"):
        # measured after 100 steps: <= 3.5e-3 (sgd, momentum), 7.8e-3 (Adam normalises the update: sign-level noise counts fully)
        assert float((sd[k].double() - p[k]).norm() / p[k].norm()) < (2e-2 if kind == "adam" else 6e-3), k
    assert losses[-1] < losses[0]
    # validation / prediction on the fabric: forward-only launches, nothing pushed, step counter untouched
    xv, yv = torch.from_numpy(xs[5000:5300]), torch.from_numpy(ys[5000:5300])
    ev = eng.evaluate(xv, yv)
    h = torch.relu(xv.double() @ p["hid_w"] + p["hid_b"])
    z = h @ p["sm_w"] + p["sm_b"]
    ref_loss = float(-(yv.double() * torch.log(torch.clamp(torch.softmax(z, -1), 1e-10, 1.0))).sum())
    assert abs(ev["loss"] - ref_loss) < (2e-2 if kind == "adam" else 2e-3) * abs(ref_loss) and ev["count"] == 300
    assert float((ev["logits"].double().cpu() - z).norm() / z.norm()) < (1e-2 if kind == "adam" else 2e-3)
    assert ev["correct"] == int((z.argmax(1) == yv.argmax(1)).sum()) or abs(ev["correct"] - int((z.argmax(1) == yv.argmax(1)).sum())) <= 1
    assert torch.equal(eng.predict(xv).cpu(), ev["logits"].argmax(1).cpu())
    assert int(eng.state_dict()["global_step"]) == steps
    l_next = eng.step(torch.from_numpy(xs[:100]).pin_memory(), torch.from_numpy(ys[:100]).pin_memory())   # training resumes
    assert np.isfinite(l_next) and int(eng.state_dict()["global_step"]) == steps + 1
    eng.close()


@pytest.mark.parametrize("kind,splits,hctas", [("sgd", 1, 1), ("adam", 1, 4), ("sgd", 4, 4), ("sgd", 1, 8)])
def test_colocated_engine_matches_oracle(kind, splits, hctas):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.parallel.fabric import Fabric
    from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    torch.cuda.set_device(0)
    opt = {"kind": kind, "lr": 0.01 if kind == "adam" else 0.002}
    eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer=opt, seed=3, f1_splits=splits, head_ctas=hctas,
                                                precision="bf16"), Fabric(1, {0: 0}))
    eng.init_params()
    p = {k: v.clone() for k, v in eng.state_dict().items() if k in ("hid_w", "hid_b", "sm_w", "sm_b")}
    xs, ys = synthetic_mnist(1000, seed=5)
    state, losses, ref_losses = {}, [], []
    for t in range(1, 7):
        x = torch.from_numpy(xs[(t - 1) * 100:t * 100])
        y = torch.from_numpy(ys[(t - 1) * 100:t * 100])
        losses.append(eng.step(x.pin_memory(), y.pin_memory()))
        ref_losses.append(_oracle_step(p, x, y, opt, state, t))
    eng.check_errors()
    sd = eng.state_dict()
    assert int(sd["global_step"]) == 6
    np.testing.assert_allclose(losses, ref_losses, rtol=2e-2)
    for k in ("hid_w", "hid_b", "sm_w", "sm_b"):
        torch.testing.assert_close(sd[k], p[k], rtol=5e-2, atol=5e-3)
    assert losses[-1] < losses[0]
    ev = eng.evaluate(torch.from_numpy(xs[600:900]), torch.from_numpy(ys[600:900]))      # bf16 engines: op-layer kernels
    assert ev["count"] == 300 and np.isfinite(ev["loss"]) and 0.0 <= ev["accuracy"] <= 1.0
    eng.close()


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_device_dataset_and_cuda_graph_replay_are_step_exact(precision):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.parallel.fabric import Fabric
    from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    torch.cuda.set_device(0)
    xs, ys = synthetic_mnist(1200, seed=6)

    def run(graph):
        eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer={"kind": "sgd", "lr": 0.002}, seed=1,
                                                    precision=precision), Fabric(1, {0: 0}))
        eng.init_params()
        eng.attach_dataset(0, xs, ys)
        eng.enqueue_local_steps(2, "dataset")
        if graph:
            eng.capture_graphs(4, "dataset")
            n = eng.replay_graphs(2)
            assert n == 2 * 4 * (eng.launches_per_worker_step("dataset") + 1)
        else:
            eng.enqueue_local_steps(8, "dataset")
        loss = eng.read_loss()
        eng.check_errors()
        sd = eng.state_dict()
        eng.close()
        return loss, sd
    l0, s0 = run(False)
    l1, s1 = run(True)
    assert int(s0["global_step"]) == 10 and int(s1["global_step"]) == 10
    assert l0 == pytest.approx(l1, rel=1e-4)          # fp32 atomics of the row-parallel head: summation order is not fixed
    torch.testing.assert_close(s0["hid_w"], s1["hid_w"], rtol=1e-4, atol=1e-5)


def test_step_stats_trace_the_whole_step_per_task():
    """Timeline on the fabric tier (reference example_in_graph.py:65-68 traces the step per task): the ps ring has one
    ``ps_apply`` row per aggregate, the worker ring four phases per ``mlp_step_kernel`` launch, in device-clock order."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import json
    from distributed_tensorflow_b200.parallel.fabric import Fabric
    from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    from distributed_tensorflow_b200.utils.timeline import Timeline
    torch.cuda.set_device(0)
    xs, ys = synthetic_mnist(1200, seed=6)
    eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer={"kind": "sgd", "lr": 0.002}, seed=1), Fabric(1, {0: 0}))
    eng.init_params()
    eng.attach_dataset(0, xs, ys)
    eng.enqueue_local_steps(5, "dataset")
    eng.evaluate(torch.from_numpy(xs[:200]), torch.from_numpy(ys[:200]))
    ev = eng.step_stats()
    eng.check_errors()
    eng.close()
    ps = [e for e in ev if e["task"] == "/job:ps/task:0"]
    wk = [e for e in ev if e["task"] == "/job:worker/task:0"]
    assert len(ps) == 5
    for step in range(5):
        mine = sorted((e for e in wk if e["name"].startswith("mlp_step/") and e["name"].endswith("[step %d]" % step)),
                      key=lambda e: e["start_us"])
        assert [e["op"].split("/")[1] for e in mine] == ["wait_token", "forward_gemm", "head_softmax_xent", "backward_gemm_push"]
        for a, b in zip(mine, mine[1:]):
            assert b["start_us"] >= a["start_us"] + a["dur_us"] - 1.0        # (epoch microseconds in a double: 0.25 us grain)
        assert sum(e["dur_us"] for e in mine) < 1000.0               # a step is tens of microseconds, not a wrapped clock
    assert any(e["name"].startswith("mlp_forward/") for e in wk)         # the forward-only (validation) launch is traced too
    trace = json.loads(Timeline(ev).generate_chrome_trace_format())
    names = {t["args"]["name"] for t in trace["traceEvents"] if t.get("ph") == "M"}
    assert any("/job:worker/task:0" in n for n in names) and any("/job:ps/task:0" in n for n in names)


def test_smoke_entry_point():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.parallel.ps_engine import smoke_step
    out = smoke_step()
    assert out["global_step"] == 4 and out["losses"][-1] < out["losses"][0] * 1.5


def test_generic_engine_resnet18_trains_on_fabric_ps():
    """ResNet-18 under the ps API (config 5): conv-as-tcgen05-GEMM workers, fused ps_apply, one GPU colocated."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.models import resnet18_init, resnet18_loss, resnet18_param_shapes
    from distributed_tensorflow_b200.ops import cuda_lib
    from distributed_tensorflow_b200.parallel.fabric import Fabric
    from distributed_tensorflow_b200.parallel.generic_engine import GenericPSEngine
    from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig
    torch.cuda.set_device(0)
    shapes = resnet18_param_shapes(10, "cifar")
    eng = GenericPSEngine(shapes, EngineConfig(colocated=True, optimizer={"kind": "momentum", "lr": 0.05, "momentum": 0.9}),
                          Fabric(1, {0: 0}))
    init = resnet18_init(10, "cifar", seed=2)
    eng.init_params(init)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(16, 16, 16, 3, generator=g).cuda()
    y = torch.eye(10)[torch.randint(0, 10, (16,), generator=g)].cuda()
    n0 = cuda_lib.launch_count()
    losses = []
    for _ in range(6):
        losses.append(float(eng.step(resnet18_loss, {0: (x, y)})[0]))
    eng.check_errors()
    sd = eng.state_dict()
    assert int(sd["global_step"]) == 6
    assert losses[-1] < losses[0]                       # same batch every step: must go down
    assert cuda_lib.launch_count() - n0 > 6 * 60        # the convolutions really ran as our GEMMs
    # first-step loss agrees with a plain fp32 PyTorch forward of the same network
    ref = float(resnet18_loss({k: v.cuda() for k, v in init.items()}, x.cpu().cuda(), y))
    assert abs(losses[0] - ref) < 0.05 * abs(ref)
    assert not torch.equal(sd["stem/conv"], init["stem/conv"])
    eng.close()


def test_native_step_prefetch_matches_plain_steps():
    """Double-buffered input staging (copy stream + prefetch of the next batch) and the CUDA-graphed native plans
    produce exactly the parameters of the same steps issued one by one."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.parallel.fabric import Fabric
    from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    torch.cuda.set_device(0)
    xs, ys = synthetic_mnist(1200, seed=9)
    hx, hy = torch.from_numpy(xs).pin_memory(), torch.from_numpy(ys).pin_memory()
    batches = [(hx[i * 100:(i + 1) * 100], hy[i * 100:(i + 1) * 100]) for i in range(12)]

    def run(prefetch):
        eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer={"kind": "momentum", "lr": 0.001, "momentum": 0.9},
                                                    seed=4, head_ctas=1, precision="bf16"), Fabric(1, {0: 0}))     # one head CTA: no fp32 atomics, bit-exact
        eng.init_params()
        losses = []
        for i in range(11):
            losses.append(eng.step(*batches[i], prefetch=batches[i + 1] if prefetch else None))
        eng.check_errors()
        sd = eng.state_dict()
        eng.close()
        return losses, sd
    l0, s0 = run(False)
    l1, s1 = run(True)
    assert int(s0["global_step"]) == 11 and int(s1["global_step"]) == 11
    np.testing.assert_allclose(l0, l1, rtol=1e-6)
    for k in ("hid_w", "hid_b", "sm_w", "sm_b"):
        torch.testing.assert_close(s0[k], s1[k], rtol=0, atol=0)
    assert l0[-1] < l0[0]


def test_symmetric_buffers_and_nvls_collectives_two_gpus():
    """Fabric collectives on symmetric VMM buffers with ONE process driving two GPUs (in-graph topology): unicast peer
    loads/stores always, multimem.ld_reduce / multimem.st when the box has NVLS.  Needs >= 2 GPUs."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ctypes
    from distributed_tensorflow_b200.ops import cuda_lib
    from distributed_tensorflow_b200.parallel.fabric import Fabric
    lib = cuda_lib.load()
    fabric = Fabric(2, {0: 0, 1: 1})
    if fabric.nvls_level() == 0:
        pytest.skip("CUDA VMM with POSIX fd export is not available here")
    n = 256 * 1024
    grads = fabric.alloc_symmetric("t_grads", n * 4)
    repl = fabric.alloc_symmetric("t_repl", n * 4)
    grads.local(1).tensor(torch.float32, 0, n).fill_(3.0)        # the worker's gradient; rank 0 (ps) contributes zeros
    torch.cuda.synchronize(1)
    with torch.cuda.device(0):
        dst = torch.zeros(n, device="cuda:0")
        src = torch.arange(n, dtype=torch.float32, device="cuda:0")
        st = torch.cuda.current_stream().cuda_stream
        peers_g = (ctypes.c_void_p * 16)(grads.peer(0, 1).ptr)
        peers_r = (ctypes.c_void_p * 16)(repl.peer(0, 1).ptr)
        modes = [None] + ([True] if grads.multicast else [])
        for use_mc in modes:
            dst.zero_()
            assert lib.dtf_fabric_reduce(grads.mc(0) if use_mc else None, peers_g, 1, dst.data_ptr(), n, 0, st) == 0
            torch.cuda.synchronize(0)
            assert bool(torch.all(dst == 3.0))
            repl.local(1).tensor(torch.float32, 0, n).zero_()
            torch.cuda.synchronize(1)
            assert lib.dtf_fabric_bcast(src.data_ptr(), repl.mc(0) if use_mc else None, peers_r, 1, n * 4, 0, st) == 0
            torch.cuda.synchronize(0)
            assert torch.equal(repl.local(1).tensor(torch.float32, 0, n).cpu(), src.cpu())
    fabric.close()
