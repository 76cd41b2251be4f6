"""Single-machine programs: ``one_gpu`` and multi-tower data parallelism (reference ``standalone.py``, S15-S20).

* ``one_gpu()``: parameters on the host, ``add`` and ``matmul`` on ``/gpu:0``
  (golden values ``[[2,3],[6,7]]`` and ``[[5],[14]]``).
* towers: variables are allocated once on ``/cpu:0`` and shared through
  ``variable_scope`` reuse; the batch is split across ``NUM_GPU`` towers; every tower computes
  its gradients; ``average_tower_grads`` averages them per variable; ONE ``apply_gradients``.
Fixes vs the reference: targets are split from the *target* tensor, tower ``i``
runs on ``/gpu:i`` when that GPU exists (CPU otherwise), ``one_gpu`` is reachable (``--one_gpu``).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distributed_tensorflow_b200 as dtf


def one_gpu():
    with dtf.device("/cpu:0"):
        w = dtf.Variable(dtf.constant([[1.0, 2.0], [4.0, 5.0]]), name="w")
        b = dtf.Variable(dtf.constant([[1.0], [2.0]]), name="b")
    with dtf.device("/gpu:0"):
        addwb = dtf.add(w, b)
        mulwb = dtf.matmul(w, b)
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        val1, val2 = sess.run([addwb, mulwb])
        print(val1)
        print(val2)
    return val1, val2


def _allocate_variable(name, shape, initializer, dtype=dtf.float32, verbose=True):
    # parameters live in host memory; every tower reads them (device transfer is implicit)
    with dtf.device('/cpu:0'):
        var = dtf.get_variable(name, shape, initializer=initializer, dtype=dtype)
    if verbose:
        print('%s: %s' % (var.op.name, var.device))
    return var


def tower(input_tensor, target_tensor, scope, dims=(), verbose=True):
    for i, d in enumerate(dims):
        with dtf.variable_scope('affine%d' % i):
            w = _allocate_variable('w', shape=[int(input_tensor.get_shape()[1]), d],
                                   initializer=dtf.truncated_normal_initializer(0, 1), verbose=verbose)
            b = _allocate_variable('b', shape=[], initializer=dtf.zeros_initializer, verbose=verbose)
        input_tensor = dtf.nn.relu(dtf.matmul(input_tensor, w) + b)
    with dtf.variable_scope('affine_last'):
        w = _allocate_variable('w', shape=[int(input_tensor.get_shape()[1]), 1],
                               initializer=dtf.constant_initializer(value=1), verbose=verbose)
        b = _allocate_variable('b', shape=[], initializer=dtf.zeros_initializer, verbose=verbose)
    y = dtf.matmul(input_tensor, w) + b
    l = dtf.reduce_mean(dtf.square(y - target_tensor))
    dtf.add_to_collection('losses', l)
    return y, l


def average_tower_grads(tower_grads, verbose=True):
    if verbose:
        print('towerGrads:')
        for idx, grads in enumerate(tower_grads):
            print('grads---tower_%d' % idx)
            for g, v in grads:
                print('\t%s\n\t%s' % (g.op.name, v.op.name))
    if len(tower_grads) == 1:
        return tower_grads[0]
    avg = []
    for grad_var_s in zip(*tower_grads):
        grads = [dtf.expand_dims(g, 0) for g, _ in grad_var_s]
        all_g = dtf.concat(grads, 0)
        avg.append((dtf.reduce_mean(all_g, 0, keep_dims=False), grad_var_s[0][1]))
    return avg


def generate_towers(NUM_GPU=2, dim_in=1, dims=None, lr=1e-2, verbose=True):
    dims = dims or []
    input_tensor = dtf.placeholder(dtf.float32, shape=[None, dim_in], name='input')
    target_tensor = dtf.placeholder(dtf.float32, shape=[None, dim_in], name='target')
    input_tensors = dtf.split(input_tensor, NUM_GPU)     # batch must divide by NUM_GPU
    target_tensors = dtf.split(target_tensor, NUM_GPU)
    tower_grads = []
    opt = dtf.train.GradientDescentOptimizer(lr)
    import torch
    ngpu = torch.cuda.device_count()
    y = loss = None
    for i in range(NUM_GPU):
        dev = '/gpu:%d' % i if i < ngpu else '/cpu:0'
        with dtf.device(dev):
            with dtf.name_scope('tower_%d' % i) as scope:
                if verbose:
                    print("tower %d device:%s" % (i, dev))
                y, loss = tower(input_tensors[i], target_tensors[i], scope, dims, verbose=verbose and i == 0)
                dtf.get_variable_scope().reuse_variables()     # later towers share the variables
                tower_grads.append(opt.compute_gradients(loss))
    apply_gradient_op = opt.apply_gradients(average_tower_grads(tower_grads, verbose), global_step=None)
    if verbose:
        print('ALL variables:')
        for v in dtf.global_variables():
            print('\t%s' % v.op.name)
    return input_tensor, target_tensor, y, loss, apply_gradient_op


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--one_gpu", action="store_true")
    ap.add_argument("--num_gpu", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10000)
    ap.add_argument("--batch_size", type=int, default=2000)
    a = ap.parse_args()
    if a.one_gpu:
        one_gpu()
        sys.exit(0)
    sess = dtf.Session()
    dim_in, dims = 2, [64, 32]
    input_tensor, target_tensor, y, loss, apply_gradient_op = generate_towers(NUM_GPU=a.num_gpu, dim_in=dim_in, dims=dims)
    sess.run(dtf.global_variables_initializer())
    inputs = np.random.rand(a.batch_size, dim_in)
    targets = inputs * 2 + 1
    feed_dict = {input_tensor: inputs, target_tensor: targets}
    tstart = time.time()
    l = None
    for i in range(a.iters):
        _, l = sess.run([apply_gradient_op, loss], feed_dict=feed_dict)
    print('%d towers: %.2fs for %d iterations, final tower loss %g' % (a.num_gpu, time.time() - tstart, a.iters, l))
