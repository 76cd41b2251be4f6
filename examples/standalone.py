"""One machine, several devices: host-resident parameters, tower-style data parallelism with gradient averaging.

Counterpart of the reference's ``standalone.py`` (S15-S20), restructured around a small ``Towers`` class:

* ``--one_gpu``: two constants on the host, ``add`` and ``matmul`` placed on ``/gpu:0`` -> ``[[2,3],[6,7]]`` and
  ``[[5],[14]]`` (the reference defines this function but never calls it).
* default: a ReLU MLP regressor (``dims`` hidden widths, a final ``[d,1]`` layer initialised to one, scalar biases)
  whose variables are created ONCE on ``/cpu:0`` and re-used by every tower through ``variable_scope`` reuse; the
  batch -- inputs AND targets (the reference splits the inputs twice) -- is split across ``--num_gpu`` towers, tower
  ``i`` runs on ``/gpu:i`` when that GPU exists and on the CPU otherwise; each tower computes its own gradients,
  they are averaged per variable (``expand_dims`` + ``concat`` + ``reduce_mean``) and applied once by SGD.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # examples/_common.py
import argparse
import time

import numpy as np
import torch

from _common import dtf


def one_gpu():
    with dtf.device("/cpu:0"):
        w = dtf.Variable(dtf.constant([[1.0, 2.0], [4.0, 5.0]]), name="w")
        b = dtf.Variable(dtf.constant([[1.0], [2.0]]), name="b")
    with dtf.device("/gpu:0"):
        total, product = dtf.add(w, b), dtf.matmul(w, b)
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        values = sess.run([total, product])
    for v in values:
        print(v)
    return values


class Towers:
    def __init__(self, num_towers, dim_in, dims, lr=1e-2, verbose=True):
        self.verbose = verbose
        self.inputs = dtf.placeholder(dtf.float32, shape=[None, dim_in], name="input")
        self.targets = dtf.placeholder(dtf.float32, shape=[None, dim_in], name="target")
        sgd = dtf.train.GradientDescentOptimizer(lr)
        gpus = torch.cuda.device_count()
        per_tower = []
        for i, (x, t) in enumerate(zip(dtf.split(self.inputs, num_towers), dtf.split(self.targets, num_towers))):
            device = "/gpu:%d" % i if i < gpus else "/cpu:0"
            with dtf.device(device), dtf.name_scope("tower_%d" % i):
                self._say("tower %d device:%s" % (i, device))
                self.output, self.loss = self._tower(x, t, dims, announce=(i == 0))
                dtf.get_variable_scope().reuse_variables()            # towers after the first share the weights
                per_tower.append(sgd.compute_gradients(self.loss))
        self.train_op = sgd.apply_gradients(self._average(per_tower), global_step=None)
        self._say("ALL variables:\n" + "\n".join("\t%s" % v.op.name for v in dtf.global_variables()))

    def _say(self, text):
        if self.verbose:
            print(text)

    def _host_variable(self, name, shape, initializer, announce):
        with dtf.device("/cpu:0"):          # parameters stay in host memory; towers read them from there
            var = dtf.get_variable(name, shape, initializer=initializer, dtype=dtf.float32)
        if announce:
            self._say("%s: %s" % (var.op.name, var.device))
        return var

    def _tower(self, x, target, dims, announce):
        h = x
        for i, width in enumerate(dims):
            with dtf.variable_scope("affine%d" % i):
                w = self._host_variable("w", [int(h.get_shape()[1]), width], dtf.truncated_normal_initializer(0, 1), announce)
                b = self._host_variable("b", [], dtf.zeros_initializer, announce)
            h = dtf.nn.relu(dtf.matmul(h, w) + b)
        with dtf.variable_scope("affine_last"):
            w = self._host_variable("w", [int(h.get_shape()[1]), 1], dtf.constant_initializer(value=1), announce)
            b = self._host_variable("b", [], dtf.zeros_initializer, announce)
        out = dtf.matmul(h, w) + b
        mse = dtf.reduce_mean(dtf.square(out - target))
        dtf.add_to_collection("losses", mse)
        return out, mse

    def _average(self, per_tower):
        if self.verbose:
            for i, pairs in enumerate(per_tower):
                print("grads---tower_%d" % i)
                for g, v in pairs:
                    print("\t%s\n\t%s" % (g.op.name, v.op.name))
        if len(per_tower) == 1:
            return per_tower[0]
        averaged = []
        for same_var in zip(*per_tower):
            stacked = dtf.concat([dtf.expand_dims(g, 0) for g, _ in same_var], 0)
            averaged.append((dtf.reduce_mean(stacked, 0, keep_dims=False), same_var[0][1]))
        return averaged


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one_gpu", action="store_true")
    ap.add_argument("--num_gpu", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10000)
    ap.add_argument("--batch_size", type=int, default=2000)
    a = ap.parse_args()
    if a.one_gpu:
        one_gpu()
        return
    model = Towers(a.num_gpu, dim_in=2, dims=[64, 32])
    xs = np.random.rand(a.batch_size, 2)
    feed = {model.inputs: xs, model.targets: xs * 2 + 1}
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        started, last = time.time(), None
        for _ in range(a.iters):
            _, last = sess.run([model.train_op, model.loss], feed_dict=feed)
    print("%d towers: %.2fs for %d iterations, final tower loss %g" % (a.num_gpu, time.time() - started, a.iters, last))


if __name__ == "__main__":
    main()
