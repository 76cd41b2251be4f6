"""``PSTrainEngine.train_loop`` on one GPU (written after the round's GPU budget was spent: first hardware run is the driver's;
the file sorts last so that a problem here cannot hide the results of the validated tiers).  CPU coverage of the same code:
``test_step_exec_host.py`` (the native loop's call sequence) and ``test_fabric_host_logic.py`` (the Python side)."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_native_train_loop_matches_synchronous_steps(precision):
    """``PSTrainEngine.train_loop`` (K steps enqueued by one native call, csrc/step_exec.cu dtf_run_loop) computes what K
    ``step()`` calls compute: same per-step losses, same parameters, same global step -- also when the loop is entered with a
    batch ``step()`` prefetched, and when it is called twice."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from distributed_tensorflow_b200.ops import cuda_lib
    from distributed_tensorflow_b200.parallel.fabric import Fabric
    from distributed_tensorflow_b200.parallel.ps_engine import EngineConfig, MLPSpec, PSTrainEngine
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    torch.cuda.set_device(0)
    xs, ys = synthetic_mnist(700, seed=11)
    hx, hy = torch.from_numpy(xs).pin_memory().view(7, 100, 784), torch.from_numpy(ys).pin_memory().view(7, 100, 10)
    K = 23

    def run(native):
        eng = PSTrainEngine(MLPSpec(), EngineConfig(colocated=True, optimizer={"kind": "momentum", "lr": 0.001, "momentum": 0.9},
                                                    seed=4, head_ctas=1, precision=precision), Fabric(1, {0: 0}))
        eng.init_params()
        if native:
            before = cuda_lib.launch_count()
            a = eng.train_loop(hx, hy, 9, first=2, stride=3, depth=2, prefetch_next=True)     # 4 eager steps + 5 in the native loop; the batch of step 9 is prefetched
            extra = eng.step(hx[(2 + 9 * 3) % 7], hy[(2 + 9 * 3) % 7], prefetch=(hx[(2 + 10 * 3) % 7], hy[(2 + 10 * 3) % 7]))
            b = eng.train_loop(hx, hy, K - 10, first=2 + 10 * 3, stride=3, depth=4)    # enters with its first batch prefetched
            losses = list(a) + [extra] + list(b)
            assert cuda_lib.launch_count() - before >= K
        else:
            losses = [eng.step(hx[(2 + 3 * i) % 7], hy[(2 + 3 * i) % 7]) for i in range(K)]
        eng.check_errors()
        sd = eng.state_dict()
        eng.close()
        return losses, sd
    l0, s0 = run(False)
    l1, s1 = run(True)
    assert len(l0) == len(l1) == K and int(s0["global_step"]) == K and int(s1["global_step"]) == K
    exact = precision == "bf16"
    np.testing.assert_allclose(l0, l1, rtol=1e-6 if exact else 5e-5)
    for k in ("hid_w", "hid_b", "sm_w", "sm_b"):
        torch.testing.assert_close(s0[k], s1[k], rtol=0 if exact else 1e-4, atol=0 if exact else 1e-5)
