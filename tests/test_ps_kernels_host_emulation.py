"""The device-side parameter-server protocol (csrc/ps_engine.cu: ps_apply, push_grad, wait_token, publish, fabric
collectives) executed on the CPU: the SAME kernel source compiled by g++ against tests/emu/host_emu.h (threads + barriers
= a thread block, blocks one after another, sequentially consistent atomics for the scoped PTX accesses, a registry of
member buffers for multimem).  Covers what SURVEY A11/A12 specify for the ps: fresh / stale decision, backup workers,
gradient MEAN, SGD / Momentum / TF-Adam, bf16 publication (unicast replicas and multicast), zero-after-read ranges,
tokens and versions, async round-robin + staleness histogram, bounded waits.  Memory-model races between concurrently
running blocks/GPUs are outside what an emulation can show: that is the hardware tier (tests/test_gpu_*.py, tools/mp_check.py).
"""
import ctypes
import math
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from distributed_tensorflow_b200.ops.cuda_lib import MAX_WORKERS, PsApplyArgs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "distributed_tensorflow_b200", "csrc")
CTL_FIELDS = ["global_step", "param_version", "beta1_power", "beta2_power", "dropped_stale", "applied_total",
              "staleness_hist", "staleness_sum", "err", "w", "w_stride", "consumed"]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libps_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-DDTF_HOST_EMU", "-I" + os.path.join(ROOT, "tests", "emu"), "-I" + CSRC,
                    "-x", "c++", "-shared", "-fPIC", "-pthread", "-o", so, os.path.join(CSRC, "ps_engine.cu")], check=True)
    lib = ctypes.CDLL(so)
    vp, ll, i, ull = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_ulonglong
    lib.dtf_ps_apply.argtypes = [ctypes.POINTER(PsApplyArgs), vp]
    lib.dtf_ps_publish.argtypes = [vp, vp, ll, vp]
    lib.dtf_wait_token.argtypes = [vp, ull, vp, ull, vp, vp]
    lib.dtf_push_grad.argtypes = [vp, vp, ll, vp, vp, i, i, i, i, vp]
    lib.dtf_pull_shadow.argtypes = [vp, vp, ll, i, vp]
    lib.dtf_fabric_bcast.argtypes = [vp, vp, vp, i, ll, i, vp]
    lib.dtf_fabric_reduce.argtypes = [vp, vp, i, vp, ll, i, vp]
    lib.dtf_offsetof_ctl.argtypes = [i]
    lib.dtf_emu_mc_register.argtypes = [vp, ll, vp, i]
    lib.off = {k: lib.dtf_offsetof_ctl(j) for j, k in enumerate(CTL_FIELDS)}
    return lib


def _aligned_bytes(nbytes, align=128):
    """Zeroed host buffer whose data pointer is ``align``-byte aligned (the control structs are __align__(128))."""
    raw = torch.zeros(nbytes + align, dtype=torch.uint8)
    off = (-raw.data_ptr()) % align
    return raw[off:off + nbytes], raw


class World:
    """Host-memory stand-in for one ps shard and its workers (what PSTrainEngine allocates on the GPUs)."""

    def __init__(self, lib, n=1280, workers=3, kind=0, sync=True, R=None, ctas_per_push=2, lr=0.1, **opt):
        self.lib, self.n, self.W, self.cpp = lib, n, workers, ctas_per_push
        self.ctl, self._k0 = _aligned_bytes(lib.dtf_sizeof_ps_control())
        self.mb_bytes = lib.dtf_sizeof_mailbox()
        self.mb, self._k1 = _aligned_bytes(self.mb_bytes * workers)
        g = torch.Generator().manual_seed(n + workers)
        self.master = torch.randn(n, generator=g)
        self.master0 = self.master.clone()
        self.slot_m, self.slot_v = torch.zeros(n), torch.zeros(n)
        self.shadow = torch.zeros(n, dtype=torch.bfloat16)
        self.grads = [torch.zeros(n) for _ in range(workers)]              # separate allocations: 16-byte aligned slots
        self.replicas = [torch.zeros(n, dtype=torch.bfloat16) for _ in range(workers)]
        a = PsApplyArgs()
        a.ctl, a.master, a.slot_m, a.slot_v = self.ctl.data_ptr(), self.master.data_ptr(), self.slot_m.data_ptr(), self.slot_v.data_ptr()
        a.shadow = self.shadow.data_ptr()
        for w in range(workers):
            a.grad[w] = self.grads[w].data_ptr()
            a.mailbox[w] = self.mb.data_ptr() + w * self.mb_bytes
        a.n, a.num_workers, a.replicas_to_aggregate, a.ctas_per_push = n, workers, R or workers, ctas_per_push
        a.mode, a.kind, a.lr = (0 if sync else 1), kind, lr
        a.momentum, a.nesterov = opt.get("momentum", 0.0), int(opt.get("nesterov", False))
        a.beta1, a.beta2, a.eps = opt.get("beta1", 0.9), opt.get("beta2", 0.999), opt.get("eps", 1e-8)
        a.timeout_ns, a.system_scope = 200_000_000, 1
        self.a = a
        self.f32("beta1_power", 2)[:] = torch.tensor([a.beta1, a.beta2])

    # -- typed views into the control block / mailboxes -------------------------------------------------
    def u64(self, field, count=1, extra=0):
        o = self.lib.off[field] + extra
        return self.ctl[o:o + 8 * count].view(torch.int64)

    def u32(self, field):
        o = self.lib.off[field]
        return self.ctl[o:o + 4].view(torch.int32)

    def f32(self, field, count=1):
        o = self.lib.off[field]
        return self.ctl[o:o + 4 * count].view(torch.float32)

    def slot(self, w):          # (arrivals, stamp) of worker w
        return self.u64("w", 2, w * self.lib.off["w_stride"])

    def mailbox(self, w):       # (token, version)
        return self.mb[w * self.mb_bytes:w * self.mb_bytes + 16].view(torch.int64)

    def push(self, w, grad, stamp):
        """What a worker's kernels do: gradient into its slot, stamp, then the arrivals of all its pushing CTAs."""
        self.grads[w].copy_(grad)
        s = self.slot(w)
        s[1] = stamp
        s[0] += self.cpp

    def apply(self):
        assert self.lib.dtf_ps_apply(ctypes.byref(self.a), None) == 0


def _bf16(t):
    return t.bfloat16()


def test_sync_sgd_mean_tokens_and_shadow(emu):
    wd = World(emu, n=1280, workers=3, kind=0, lr=0.1)
    gs = [torch.randn(1280) for _ in range(3)]
    for w in range(3):
        wd.push(w, gs[w], stamp=0)
    wd.apply()
    want = wd.master0 - 0.1 * (gs[0] + gs[1] + gs[2]) / 3
    torch.testing.assert_close(wd.master, want, rtol=1e-6, atol=1e-6)
    assert torch.equal(wd.shadow, _bf16(wd.master))
    assert int(wd.u64("global_step")) == 1 and int(wd.u64("param_version")) == 1 and int(wd.u64("applied_total")) == 3
    for w in range(3):
        assert wd.mailbox(w).tolist() == [1, 1]                 # token and version carry the NEW global step
        assert int(wd.u64("consumed", 1, 8 * w)) == wd.cpp
    assert int(wd.u32("err")) == 0
    # second aggregate: stamps must be >= the current global step (1) to count
    for w in range(3):
        wd.push(w, gs[w], stamp=1)
    wd.apply()
    torch.testing.assert_close(wd.master, want - 0.1 * (gs[0] + gs[1] + gs[2]) / 3, rtol=1e-6, atol=1e-6)
    assert int(wd.u64("global_step")) == 2 and wd.mailbox(2).tolist() == [2, 2]


def test_incomplete_push_is_not_consumed_until_all_its_ctas_arrived(emu):
    wd = World(emu, workers=2, ctas_per_push=4)
    wd.a.timeout_ns = 2_000_000
    wd.push(0, torch.ones(wd.n), 0)
    wd.grads[1].fill_(1.0)
    wd.slot(1)[0] += 3                                           # 3 of 4 CTAs of worker 1 have arrived
    wd.apply()
    assert int(wd.u32("err")) == 2 and int(wd.u64("global_step")) == 0 and torch.equal(wd.master, wd.master0)
    wd.u32("err")[0] = 0
    wd.slot(1)[0] += 1                                           # the last CTA
    wd.apply()
    assert int(wd.u32("err")) == 0 and int(wd.u64("global_step")) == 1
    torch.testing.assert_close(wd.master, wd.master0 - 0.1 * torch.ones(wd.n))


def test_backup_workers_and_stale_drop(emu):
    """replicas_to_aggregate=2 of 3: the aggregate goes ahead with two fresh gradients; the straggler's push, stamped
    with the old step, is dropped by the next aggregate, which then waits for fresh ones."""
    wd = World(emu, workers=3, R=2, lr=0.5)
    g = [torch.full((wd.n,), float(w + 1)) for w in range(3)]
    wd.push(0, g[0], 0)
    wd.push(1, g[1], 0)
    wd.apply()
    torch.testing.assert_close(wd.master, wd.master0 - 0.5 * (g[0] + g[1]) / 2)
    assert int(wd.u64("global_step")) == 1 and int(wd.u64("dropped_stale")) == 0
    assert [wd.mailbox(w)[0].item() for w in range(3)] == [1, 1, 1]        # every replica gets a token, pushers or not
    m1 = wd.master.clone()
    wd.push(2, g[2], 0)          # straggler: computed against step 0
    wd.push(0, g[0], 1)
    wd.push(1, g[1], 1)
    wd.apply()
    assert int(wd.u64("dropped_stale")) == 1 and int(wd.u64("global_step")) == 2
    torch.testing.assert_close(wd.master, m1 - 0.5 * (g[0] + g[1]) / 2)    # the stale gradient did not contribute
    assert int(wd.u64("consumed", 1, 16)) == wd.cpp                          # ...but its arrivals were consumed


def test_timeout_sets_error_unless_idle_polling(emu):
    wd = World(emu, workers=2)
    wd.a.timeout_ns = 1_000_000
    wd.a.idle_ok = 1
    wd.apply()
    assert int(wd.u32("err")) == 0 and int(wd.u64("global_step")) == 0 and int(wd.u64("param_version")) == 1
    wd.a.idle_ok = 0
    wd.apply()
    assert int(wd.u32("err")) == 2 and wd.mailbox(0).tolist() == [0, 0]


def test_async_applies_each_push_alone_round_robin_with_staleness(emu):
    wd = World(emu, workers=3, sync=False, lr=0.1)
    g = [torch.full((wd.n,), float(w + 1)) for w in range(3)]
    for w in (2, 0):                     # two pushes pending, both computed from version 0
        wd.push(w, g[w], stamp=0)
    wd.apply()                           # round robin starts after last_async_worker (0) -> worker 2? no: 1 is empty -> 2
    wd.apply()
    torch.testing.assert_close(wd.master, wd.master0 - 0.1 * g[2] - 0.1 * g[0])       # no mean: each push applied alone
    assert int(wd.u64("global_step")) == 2
    hist = wd.u64("staleness_hist", 16).tolist()
    assert hist[0] == 1 and hist[1] == 1 and sum(hist) == 2 and int(wd.u64("staleness_sum")) == 1
    assert wd.mailbox(2).tolist() == [1, 1] and wd.mailbox(0).tolist() == [1, 2]      # acks only to the pusher; version = gs
    assert wd.mailbox(1).tolist() == [0, 0]


def test_momentum_and_nesterov(emu):
    for nesterov in (False, True):
        wd = World(emu, workers=1, kind=1, lr=0.1, momentum=0.9, nesterov=nesterov)
        p, acc = wd.master0.clone(), torch.zeros(wd.n)
        for t in range(3):
            g = torch.randn(wd.n, generator=torch.Generator().manual_seed(t))
            wd.push(0, g, t)
            wd.apply()
            acc = 0.9 * acc + g
            p = p - (0.1 * g + 0.1 * 0.9 * acc if nesterov else 0.1 * acc)
        torch.testing.assert_close(wd.master, p, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(wd.slot_m, acc, rtol=1e-5, atol=1e-5)


def test_tf_adam_formulation_and_beta_powers(emu):
    wd = World(emu, workers=2, kind=2, lr=0.01)
    p, m, v = wd.master0.double(), torch.zeros(wd.n).double(), torch.zeros(wd.n).double()
    for t in range(1, 4):
        gs = [torch.randn(wd.n, generator=torch.Generator().manual_seed(10 * t + w)) for w in range(2)]
        for w in range(2):
            wd.push(w, gs[w], t - 1)
        wd.apply()
        g = ((gs[0] + gs[1]) / 2).double()
        lr_t = 0.01 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)        # epsilon OUTSIDE the bias correction (TF)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        p = p - lr_t * m / (v.sqrt() + 1e-8)
    torch.testing.assert_close(wd.master.double(), p, rtol=2e-4, atol=2e-5)
    b = wd.f32("beta1_power", 2)
    assert abs(float(b[0]) - 0.9 ** 4) < 1e-6 and abs(float(b[1]) - 0.999 ** 4) < 1e-6


def test_zero_after_read_ranges_and_unicast_replica_publish(emu):
    wd = World(emu, workers=2)
    wd.a.publish_replicas = 1
    for w in range(2):
        wd.a.replica[w] = wd.replicas[w].data_ptr()
    wd.a.num_zero = 1
    wd.a.zero_begin[0], wd.a.zero_end[0] = 64, 130              # not float4-aligned at the end on purpose
    for w in range(2):
        wd.push(w, torch.ones(wd.n), 0)
    wd.apply()
    for w in range(2):
        assert torch.equal(wd.replicas[w], _bf16(wd.master))
        assert float(wd.grads[w][64:130].abs().sum()) == 0.0 and float(wd.grads[w][:64].min()) == 1.0
        assert float(wd.grads[w][130:].min()) == 1.0


@pytest.mark.parametrize("n", [1280, 1030])
def test_multicast_reduce_and_publish_match_the_unicast_path(emu, n):
    """NVLS mode: gradients stay in the workers' copies of a symmetric buffer (the ps's own copy holds zeros),
    multimem.ld_reduce sums them, ONE multimem.st updates every replica, zero-after-read clears every copy."""
    W = 3
    uni, mc = World(emu, n=n, workers=W), World(emu, n=n, workers=W)
    g = [torch.randn(n, generator=torch.Generator().manual_seed(w)) for w in range(W)]
    ps_copy = torch.zeros(n)                                       # the ps GPU's own copy of the symmetric gradient buffer
    fake_g, fake_r = torch.zeros(n), torch.zeros(n, dtype=torch.bfloat16)          # placeholder "multicast" address ranges
    members_g = (ctypes.c_void_p * (W + 1))(ps_copy.data_ptr(), *[mc.grads[w].data_ptr() for w in range(W)])
    ps_repl = torch.zeros(n, dtype=torch.bfloat16)
    members_r = (ctypes.c_void_p * (W + 1))(ps_repl.data_ptr(), *[mc.replicas[w].data_ptr() for w in range(W)])
    emu.dtf_emu_mc_register(fake_g.data_ptr(), n * 4, members_g, W + 1)
    emu.dtf_emu_mc_register(fake_r.data_ptr(), n * 2, members_r, W + 1)
    mc.a.grad_mc, mc.a.shadow_mc = fake_g.data_ptr(), fake_r.data_ptr()
    mc.a.shadow = ps_repl.data_ptr()
    for wd in (uni, mc):
        wd.a.num_zero = 1
        wd.a.zero_begin[0], wd.a.zero_end[0] = 128, 256
        for w in range(W):
            wd.push(w, g[w], 0)
        wd.apply()
    torch.testing.assert_close(mc.master, uni.master, rtol=1e-6, atol=1e-6)
    nv = n // 4 * 4                                               # the scalar tail is written through `shadow` only
    for w in range(W):
        assert torch.equal(mc.replicas[w][:nv], _bf16(mc.master)[:nv])
        assert float(mc.grads[w][128:256].abs().sum()) == 0.0 and float(mc.grads[w][:128].abs().sum()) > 0
    assert torch.equal(ps_repl, _bf16(mc.master))
    # a missing push (backup workers) falls back to unicast reads of the chosen workers' copies
    mc.a.replicas_to_aggregate = 2
    m1 = mc.master.clone()
    mc.push(0, g[0], 1)
    mc.push(2, g[2], 1)
    mc.apply()
    want = m1 - 0.1 * (g[0] + g[2]) / 2
    want[128:256] = m1[128:256] - 0.1 * (g[0] + g[2])[128:256] / 2
    torch.testing.assert_close(mc.master, want, rtol=1e-6, atol=1e-6)
    emu.dtf_emu_mc_clear()


def test_push_grad_wait_token_publish_and_pull(emu):
    wd = World(emu, n=1030, workers=2)
    src = torch.randn(1030)
    wd.mailbox(1)[0], wd.mailbox(1)[1] = 7, 9                    # token 7, version 9
    mbp = wd.mb.data_ptr() + wd.mb_bytes
    slot1 = torch.zeros(1030)                      # 16-byte aligned like the engine's slots (1030: exercises the scalar tail)
    assert emu.dtf_push_grad(src.data_ptr(), slot1.data_ptr(), 1030, wd.ctl.data_ptr(), mbp, 1, 0, 1, 3, None) == 0
    assert torch.equal(slot1, src) and wd.slot(1).tolist() == [3, 7]                  # 3 CTAs arrived, stamp = token (sync)
    assert emu.dtf_push_grad(None, None, 0, wd.ctl.data_ptr(), mbp, 1, 1, 1, 1, None) == 0     # signal-only, async stamp
    assert wd.slot(1).tolist() == [4, 9]
    err = torch.zeros(1, dtype=torch.int32)
    assert emu.dtf_wait_token(mbp, 7, None, 1_000_000, err.data_ptr(), None) == 0 and int(err) == 0
    base = torch.tensor([3], dtype=torch.int64)                                        # target = 5 + *ptr = 8 > token 7
    assert emu.dtf_wait_token(mbp, 5, base.data_ptr(), 1_000_000, err.data_ptr(), None) == 0 and int(err) == 3
    shadow = torch.zeros(1030, dtype=torch.bfloat16)
    assert emu.dtf_ps_publish(wd.master.data_ptr(), shadow.data_ptr(), 1030, None) == 0
    assert torch.equal(shadow, _bf16(wd.master))
    dst = torch.zeros(1024, dtype=torch.bfloat16)
    assert emu.dtf_pull_shadow(shadow.data_ptr(), dst.data_ptr(), 2048, 2, None) == 0
    assert torch.equal(dst, shadow[:1024])


def test_fabric_broadcast_and_reduce_unicast_and_multicast(emu):
    n = 4096 + 8
    src = torch.arange(n, dtype=torch.float32)
    copies = [torch.zeros(n) for _ in range(3)]
    peers = (ctypes.c_void_p * MAX_WORKERS)(*[c.data_ptr() for c in copies])
    assert emu.dtf_fabric_bcast(src.data_ptr(), None, peers, 3, n * 4, 2, None) == 0
    assert all(torch.equal(c, src) for c in copies)
    out = torch.zeros(n)
    assert emu.dtf_fabric_reduce(None, peers, 3, out.data_ptr(), n, 2, None) == 0
    assert torch.equal(out, 3 * src)
    fake = torch.zeros(n)
    emu.dtf_emu_mc_register(fake.data_ptr(), n * 4, peers, 3)
    for c in copies:
        c.zero_()
    assert emu.dtf_fabric_bcast(src.data_ptr(), fake.data_ptr(), peers, 0, n * 4, 3, None) == 0
    assert all(torch.equal(c, src) for c in copies)
    out.zero_()
    assert emu.dtf_fabric_reduce(fake.data_ptr(), peers, 0, out.data_ptr(), n, 3, None) == 0
    assert torch.equal(out, 3 * src)
    assert emu.dtf_fabric_reduce(None, peers, 3, out.data_ptr(), n - 2, 1, None) == -2        # not a multiple of 4 floats
    emu.dtf_emu_mc_clear()


@pytest.mark.parametrize("ctas,split_k", [(1, False), (8, False), (4, True)])
def test_fused_mlp_head_matches_pytorch(emu, ctas, split_k):
    """mlp_head_kernel (K2 + K3 + the small parts of K4, fused with the push of dW2 / db2 / db1 and the arrival signal):
    logits, softmax, the reference's clipped batch-SUM cross-entropy, dlogits with the clip gate, dW2, db2, dh (bf16, feeds
    the dW1 GEMM), db1 -- one CTA (plain stores) and row-parallel CTAs (fp32 atomics into the zeroed ps slot)."""
    from distributed_tensorflow_b200.ops.cuda_lib import MlpHeadArgs
    emu.dtf_mlp_head.argtypes = [ctypes.POINTER(MlpHeadArgs), ctypes.c_void_p]
    B, H, C, ldh, ldw = 100, 100, 10, 104, 16
    g = torch.Generator().manual_seed(ctas)
    b1 = torch.randn(H, generator=g) * 0.1
    acc = torch.randn(B, ldh, generator=g)
    acc[:, H:] = 0
    hf = torch.relu(acc[:, :H] + b1) if split_k else torch.relu(torch.randn(B, H, generator=g))
    h16 = torch.zeros(B, ldh, dtype=torch.bfloat16)
    h16[:, :H] = hf.bfloat16()
    hq = hf if split_k else h16[:, :H].float()              # what the kernel sees: fp32 (split-K) or bf16-rounded h
    w2 = torch.zeros(H, ldw, dtype=torch.bfloat16)
    w2[:, :C] = (torch.randn(H, C, generator=g) * 0.3).bfloat16()
    b2 = torch.randn(C, generator=g) * 0.1
    labels = torch.nn.functional.one_hot(torch.randint(0, C, (B,), generator=g), C).float()
    hq[0] = 0.0                                              # row 0: logits = b2; push class 3 far below the clip
    if split_k:
        acc[0] = -10.0
    else:
        h16[0] = 0
    b2c = b2.clone()
    b2c[3] = -70.0
    labels[0] = 0.0
    labels[0, 3] = 1.0

    wd = World(emu, workers=2)
    wd.mailbox(1)[0], wd.mailbox(1)[1] = 5, 6
    loss = torch.zeros(16)
    stepctr = torch.tensor([41], dtype=torch.int64)
    dh = torch.full((B, ldh), 7.0, dtype=torch.bfloat16)
    gw2, gb2, gb1 = torch.zeros(H, ldw), torch.zeros(C), torch.zeros(H)
    logits = torch.zeros(B, C)
    a = MlpHeadArgs()
    a.h, a.ldh, a.w2, a.ldw2, a.b2 = h16.data_ptr(), ldh, w2.data_ptr(), ldw, b2c.data_ptr()
    a.labels, a.ldl, a.B, a.H, a.C, a.clip_min = labels.data_ptr(), C, B, H, C, 1e-10
    a.loss_out, a.step_counter = loss.data_ptr(), stepctr.data_ptr()
    a.dh, a.lddh, a.gw2, a.ldgw2, a.gb2, a.gb1 = dh.data_ptr(), ldh, gw2.data_ptr(), ldw, gb2.data_ptr(), gb1.data_ptr()
    a.logits_out = logits.data_ptr()
    a.mailbox, a.ctl, a.rank, a.stamp_from_version = wd.mb.data_ptr() + wd.mb_bytes, wd.ctl.data_ptr(), 1, 0
    a.sys_scope, a.ctas = 1, ctas
    if split_k:
        a.h_acc, a.ld_acc, a.b1 = acc.data_ptr(), ldh, b1.data_ptr()
    assert emu.dtf_mlp_head(ctypes.byref(a), None) == 0

    W = w2[:, :C].float()
    z = (hq.double() @ W.double() + b2c.double()).requires_grad_()
    y = torch.softmax(z, -1)
    want_loss = -(labels.double() * torch.log(torch.clamp(y, 1e-10, 1.0))).sum()
    (dl,) = torch.autograd.grad(want_loss, z)
    n_ctas = (B + ((B + ctas - 1) // ctas + 7) // 8 * 8 - 1) // (((B + ctas - 1) // ctas + 7) // 8 * 8)
    torch.testing.assert_close(loss[:n_ctas].double().sum(), want_loss.detach(), rtol=1e-5, atol=1e-3)
    assert float(loss[n_ctas:].abs().sum()) == 0.0
    torch.testing.assert_close(logits.double(), z.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gw2[:, :C].double(), hq.double().t() @ dl, rtol=1e-4, atol=1e-4)
    assert float(gw2[:, C:].abs().sum()) == 0.0
    torch.testing.assert_close(gb2.double(), dl.sum(0), rtol=1e-4, atol=1e-4)
    dhf = (dl @ W.double().t()) * (hq.double() > 0)
    torch.testing.assert_close(gb1.double(), dhf.sum(0), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dh[:, :H].float().double(), dhf, rtol=1e-2, atol=1e-3)           # bf16 output
    assert float(dl[0].abs().max()) == 0.0 and float(dh[0, :H].float().abs().max()) == 0.0       # the clipped row
    assert int(stepctr) == 42                                                                      # device step counter
    assert wd.slot(1).tolist() == [n_ctas, 5]                      # one arrival per CTA, stamp = the token (sync)
    if split_k:
        assert float(acc[:, :H].abs().sum()) == 0.0               # accumulator cleared for the next step


@pytest.mark.parametrize("sync", [True, False])
@pytest.mark.parametrize("ps_on_workers", [False, True])
def test_concurrent_protocol_simulation(emu, sync, ps_on_workers):
    """Liveness + result of the whole protocol with REAL concurrency: one host thread per emulated GPU stream runs the
    kernels a rank enqueues -- workers: wait_token -> (gradient) -> push_grad; ps: ps_apply -- against shared host memory.
    ``ps_on_workers``: the ps shard shares worker 0's stream, i.e. thread 0 runs [wait, push, apply...] per step
    (EngineConfig.ps_on_workers); otherwise a dedicated ps thread.  Sync: K aggregates of the mean of W gradients;
    async: W*K single-push applies.  No thread may time out, every token must arrive, the parameters must match."""
    import threading
    W, K, n, lr = 3, 5, 1024, 0.1
    wd = World(emu, n=n, workers=W, sync=sync, ctas_per_push=1, lr=lr)
    wd.a.timeout_ns = 20_000_000_000
    errs = [torch.zeros(1, dtype=torch.int32) for _ in range(W)]
    srcs = [torch.zeros(n) for _ in range(W)]
    failures = []

    def grad(w, t):
        return torch.full((n,), float(1 + w + 10 * t))

    def apply_once():
        assert emu.dtf_ps_apply(ctypes.byref(wd.a), None) == 0

    def worker(w):
        try:
            mbp = wd.mb.data_ptr() + w * wd.mb_bytes
            for t in range(K):
                assert emu.dtf_wait_token(mbp, t, None, 20_000_000_000, errs[w].data_ptr(), None) == 0
                assert int(errs[w]) == 0, "worker %d timed out waiting for token %d" % (w, t)
                srcs[w].copy_(grad(w, t))
                assert emu.dtf_push_grad(srcs[w].data_ptr(), wd.grads[w].data_ptr(), n, wd.ctl.data_ptr(), mbp, w,
                                         0 if sync else 1, 1, 1, None) == 0
                if ps_on_workers and w == 0:
                    for _ in range(1 if sync else W):
                        apply_once()
        except BaseException as e:      # noqa: BLE001
            failures.append(e)

    def ps():
        try:
            for _ in range(K if sync else K * W):
                apply_once()
        except BaseException as e:      # noqa: BLE001
            failures.append(e)

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(W)]
    if not ps_on_workers:
        threads.append(threading.Thread(target=ps))
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not failures, failures
    assert not any(t.is_alive() for t in threads)
    assert int(wd.u32("err")) == 0
    total = sum(grad(w, t) for w in range(W) for t in range(K))
    if sync:
        assert int(wd.u64("global_step")) == K and int(wd.u64("dropped_stale")) == 0
        torch.testing.assert_close(wd.master, wd.master0 - lr * total / W, rtol=1e-5, atol=1e-4)
        assert [int(wd.mailbox(w)[0]) for w in range(W)] == [K] * W
    else:
        assert int(wd.u64("global_step")) == K * W and sum(wd.u64("staleness_hist", 16).tolist()) == K * W
        torch.testing.assert_close(wd.master, wd.master0 - lr * total, rtol=1e-5, atol=1e-4)      # every push applied alone
        assert [int(wd.mailbox(w)[0]) for w in range(W)] == [K] * W                                # one ack per push


def test_concurrent_backup_worker_with_a_straggler(emu):
    """replicas_to_aggregate = 3 of 4 with one slow replica, real concurrency: the job never waits for the straggler, its
    late pushes are either aggregated (stamped with the then-current step) or dropped as stale, every replica keeps
    receiving tokens, and the counters add up."""
    import threading
    import time
    W, R, K, n = 4, 3, 8, 1024
    wd = World(emu, n=n, workers=W, R=R, ctas_per_push=1, lr=0.01)
    wd.a.timeout_ns = 20_000_000_000
    errs = [torch.zeros(1, dtype=torch.int32) for _ in range(W)]
    srcs = [torch.ones(n) for _ in range(W)]
    failures = []

    def worker(w):
        try:
            mbp = wd.mb.data_ptr() + w * wd.mb_bytes
            for t in range(K):
                assert emu.dtf_wait_token(mbp, t if w < 3 else 0, None, 20_000_000_000, errs[w].data_ptr(), None) == 0
                assert int(errs[w]) == 0
                if w == 3:
                    time.sleep(0.03)                     # the straggler
                assert emu.dtf_push_grad(srcs[w].data_ptr(), wd.grads[w].data_ptr(), n, wd.ctl.data_ptr(), mbp, w, 0, 1, 1, None) == 0
        except BaseException as e:      # noqa: BLE001
            failures.append(e)

    def ps():
        try:
            for _ in range(K):
                assert emu.dtf_ps_apply(ctypes.byref(wd.a), None) == 0
        except BaseException as e:      # noqa: BLE001
            failures.append(e)

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(W)] + [threading.Thread(target=ps)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not failures, failures
    assert not any(t.is_alive() for t in threads)
    assert int(wd.u32("err")) == 0 and int(wd.u64("global_step")) == K
    applied, dropped = int(wd.u64("applied_total")), int(wd.u64("dropped_stale"))
    assert 3 * K <= applied <= 4 * K and 0 <= dropped <= K and applied + dropped <= 4 * K
    assert all(int(wd.mailbox(w)[0]) == K for w in range(W))              # the straggler got every token too
    # all gradients are ones: every aggregate is a mean of ones whatever its size, so the result is exact
    torch.testing.assert_close(wd.master, wd.master0 - 0.01 * K, rtol=1e-5, atol=1e-5)


def test_stage_from_dataset_walks_batches_by_device_step_counter(emu):
    """Input pipeline stage of the device-resident dataset: batch index = (step * stride + offset) % nbatches with the step
    read from the worker's device counter (CUDA-graph replayable), fp32 images -> bf16 staging tile, labels copied."""
    vp, ll, i = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    emu.dtf_stage_from_dataset.argtypes = [vp, vp, ll, i, i, i, ll, ll, vp, vp, vp, vp]
    nb, B, D, C = 5, 8, 16, 4
    g = torch.Generator().manual_seed(3)
    images, labels = torch.randn(nb, B, D, generator=g), torch.rand(nb, B, C, generator=g)
    x16, lab = torch.zeros(B, D, dtype=torch.bfloat16), torch.zeros(B, C)
    for step, stride, offset in [(0, 3, 1), (4, 3, 1), (7, 1, 0)]:
        ctr = torch.tensor([step], dtype=torch.int64)
        assert emu.dtf_stage_from_dataset(images.data_ptr(), labels.data_ptr(), nb, B, D, C, stride, offset, ctr.data_ptr(),
                                          x16.data_ptr(), lab.data_ptr(), None) == 0
        bi = (step * stride + offset) % nb
        assert torch.equal(x16, images[bi].bfloat16()) and torch.equal(lab, labels[bi])
    assert emu.dtf_stage_from_dataset(images.data_ptr(), labels.data_ptr(), nb, 3, 5, C, 1, 0, None, x16.data_ptr(), lab.data_ptr(), None) == -2


def test_ps_apply_trace_ring_feeds_the_timeline(emu):
    """Tracing on the fabric tier: every ps_apply launch stamps (kind, start ns, end ns, global step) into a
    ring in ps memory; ``events_from_ring`` + ``Timeline`` turn it into a chrome trace with one process per ps GPU."""
    import json
    from distributed_tensorflow_b200.utils.timeline import Timeline, events_from_ring
    wd = World(emu, workers=2)
    ring = torch.zeros(8 * 4, dtype=torch.int64)
    wd.a.trace, wd.a.trace_cap = ring.data_ptr(), 8
    for t in range(3):
        for w in range(2):
            wd.push(w, torch.ones(wd.n), t)
        wd.apply()
    rows = [r for r in ring.view(-1, 4).tolist() if r[0] != 0]
    assert [r[3] for r in rows] == [1, 2, 3] and all(r[2] >= r[1] > 0 for r in rows)
    events = events_from_ring("/job:ps/task:0", rows, {1: "ps_apply"}, gpu_index=0)
    trace = json.loads(Timeline(events).generate_chrome_trace_format())
    names = [e["name"] for e in trace["traceEvents"] if e.get("ph") == "X"]
    assert names == ["ps_apply[step 1]", "ps_apply[step 2]", "ps_apply[step 3]"] or len(names) == len([r for r in rows if r[2] > r[1]])


def test_fp32_parameter_replicas_published_through_multicast(emu):
    """Generic models consume fp32 parameters: with ``master_mc`` the ps publishes the updated master through one
    multimem.st stream into every GPU's fp32 replica (GenericPSEngine, NVLS mode)."""
    n, W = 1024, 2
    wd = World(emu, n=n, workers=W)
    replicas = [torch.zeros(n) for _ in range(W + 1)]
    fake = torch.zeros(n)
    members = (ctypes.c_void_p * (W + 1))(*[r.data_ptr() for r in replicas])
    emu.dtf_emu_mc_register(fake.data_ptr(), n * 4, members, W + 1)
    wd.a.master_mc = fake.data_ptr()
    for w in range(W):
        wd.push(w, torch.full((n,), float(w + 1)), 0)
    wd.apply()
    torch.testing.assert_close(wd.master, wd.master0 - 0.1 * 1.5)
    for r in replicas:
        assert torch.equal(r, wd.master)
    emu.dtf_emu_mc_clear()
