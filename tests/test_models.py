"""Model families on CPU: ResNet-18 shapes / gradients, graph-API ResNet, MNIST MLP builder."""
import numpy as np
import torch

import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.models import (build_mnist_mlp, build_resnet18_graph, resnet18_init, resnet18_loss,
                                                resnet18_param_shapes)


def test_resnet18_param_count_and_forward_backward_cpu():
    shapes = resnet18_param_shapes(10, "cifar")
    n = sum(int(np.prod(s)) for _, s in shapes)
    assert 11.1e6 < n < 11.3e6                      # ResNet-18 (CIFAR stem, 10 classes) ~ 11.17 M parameters
    p = {k: v.requires_grad_(True) for k, v in resnet18_init(10, "cifar", seed=1).items()}
    x = torch.randn(4, 16, 16, 3)
    y = torch.eye(10)[torch.tensor([1, 2, 3, 4])]
    loss = resnet18_loss(p, x, y)
    grads = torch.autograd.grad(loss, list(p.values()))
    assert torch.isfinite(loss) and all(torch.isfinite(g).all() for g in grads)
    assert abs(float(loss) - np.log(10)) < 1.5


def test_resnet18_graph_api_trains_one_step():
    x = dtf.placeholder(dtf.float32, [None, 8, 8, 3])
    y_ = dtf.placeholder(dtf.float32, [None, 10])
    gs = dtf.train.get_or_create_global_step()
    logits, loss, params = build_resnet18_graph(x, y_, 10, "cifar")
    assert [v.var_name for v in dtf.trainable_variables()] == [n for n, _ in resnet18_param_shapes(10, "cifar")]
    train = dtf.train.MomentumOptimizer(0.05, 0.9).minimize(loss, global_step=gs)
    rng = np.random.RandomState(0)
    bx = rng.rand(4, 8, 8, 3).astype(np.float32)
    by = np.eye(10, dtype=np.float32)[[0, 1, 2, 3]]
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        l0 = sess.run(loss, {x: bx, y_: by})
        for _ in range(5):
            sess.run(train, {x: bx, y_: by})
        l1 = sess.run(loss, {x: bx, y_: by})
    assert l1 < l0 and sess is not None


def test_mnist_mlp_builder_fused_equals_composed():
    m1 = build_mnist_mlp(hidden=16, seed=3)
    rng = np.random.RandomState(0)
    bx, by = rng.rand(5, 784).astype(np.float32), np.eye(10, dtype=np.float32)[rng.randint(0, 10, 5)]
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        a = sess.run(m1["loss"], {m1["x"]: bx, m1["y_"]: by})
        fused = dtf.nn.clipped_softmax_xent_sum(m1["logits"], m1["y_"])
        b = sess.run(fused, {m1["x"]: bx, m1["y_"]: by})
    assert abs(a - b) < 1e-4 * abs(a)
