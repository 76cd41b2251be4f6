"""Deferred loss read-back (``step(..., sync_loss="deferred")`` -> ``PendingLoss``) on one GPU: the pipelined loop -- step
t+1 enqueued before step t's loss is waited for -- computes exactly what the synchronous loop computes.

The host logic is also covered on CPU (``test_fabric_host_logic.py::test_deferred_loss_step_ordering_with_fake_plans``); on
hardware since round 2, both precisions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_deferred_loss_pipeline_matches_synchronous_steps(precision):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.parallel.fabric import Fabric
    from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PendingLoss, PSTrainEngine
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    torch.cuda.set_device(0)
    xs, ys = synthetic_mnist(1400, seed=9)
    hx, hy = torch.from_numpy(xs).pin_memory(), torch.from_numpy(ys).pin_memory()
    batches = [(hx[i * 100:(i + 1) * 100], hy[i * 100:(i + 1) * 100]) for i in range(14)]

    def run(deferred):
        eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer={"kind": "momentum", "lr": 0.001, "momentum": 0.9},
                                                    seed=4, head_ctas=1, precision=precision), Fabric(1, {0: 0}))
        # bf16 with ONE head CTA is bit-exact; the tf32 step kernel sums its CTAs' dW2 / db partials with fp32 atomics
        eng.init_params()
        losses, pending = [], None
        for i in range(13):
            if deferred:
                h = eng.step(*batches[i], sync_loss="deferred", prefetch=batches[i + 1])
                assert isinstance(h, PendingLoss)
                if pending is not None:
                    losses.append(pending.result())
                pending = h
            else:
                losses.append(eng.step(*batches[i], prefetch=batches[i + 1]))
        if pending is not None:
            losses.append(float(pending))
            assert pending.done()
        eng.check_errors()
        sd = eng.state_dict()
        eng.close()
        return losses, sd
    l0, s0 = run(False)
    l1, s1 = run(True)
    assert len(l0) == len(l1) == 13 and int(s0["global_step"]) == 13 and int(s1["global_step"]) == 13
    exact = precision == "bf16"
    np.testing.assert_allclose(l0, l1, rtol=1e-6 if exact else 5e-5)
    for k in ("hid_w", "hid_b", "sm_w", "sm_b"):
        torch.testing.assert_close(s0[k], s1[k], rtol=0 if exact else 1e-4, atol=0 if exact else 1e-5)

