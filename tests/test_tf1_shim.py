"""The TF-1.x op shim (tools/tf1_shim.py) that lets the reference's own sources run here (golden generation only).

Checks the semantics of the ops it restates against TensorFlow's documented behaviour / independent libraries, and -
where /root/reference is present (the build container; the GPU box has none) - that the committed golden files are
what the reference's sources produce today."""
import importlib
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.linalg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
tf = importlib.import_module("tf1_shim")


def run(node, feed=None):
    return tf.Session().run(node, feed or {})


def test_fill_triangular_matches_the_tf_documentation_example():
    # tf.contrib.distributions.fill_triangular([1, 2, 3, 4, 5, 6]) == [[4, 0, 0], [6, 5, 0], [3, 2, 1]]
    got = run(tf.fill_triangular(tf.constant([1., 2., 3., 4., 5., 6.], dtype=tf.float64)))
    assert np.array_equal(got, np.array([[4., 0, 0], [6, 5, 0], [3, 2, 1]]))


def test_linear_algebra_ops_against_scipy():
    rs = np.random.RandomState(0)
    a = rs.randn(7, 7)
    k = a @ a.T + 7 * np.eye(7)
    b = rs.randn(7, 3)
    K, B = tf.placeholder(tf.float64, [None, None]), tf.placeholder(tf.float64, [None, None])
    L = tf.cholesky(K)
    sol = tf.matrix_triangular_solve(L, B, lower=True)
    l, s = tf.Session().run([L, sol], {K: k, B: b})
    l_ref = scipy.linalg.cholesky(k, lower=True)
    assert np.allclose(l, l_ref, rtol=1e-13, atol=1e-13)
    assert np.allclose(s, scipy.linalg.solve_triangular(l_ref, b, lower=True), rtol=1e-12, atol=1e-13)
    assert np.allclose(run(tf.matmul(K, B, transpose_a=True), {K: k, B: b}), k.T @ b)
    assert np.allclose(run(tf.reduce_sum(tf.log(tf.matrix_diag_part(L))), {K: k}), np.log(np.diag(l_ref)).sum())


def test_sparse_mean_pooling_matrix():
    # base.py:24-43 builds a [T, S] matrix with 1/size entries; sparse_tensor_dense_matmul = per-dataset mean
    feats = np.arange(12.0).reshape(6, 2)
    idx = np.array([[0, 0], [0, 1], [1, 2], [1, 3], [1, 4], [1, 5]])
    vals = [0.5, 0.5, 0.25, 0.25, 0.25, 0.25]
    sp = tf.SparseTensor(tf.constant(idx, dtype=tf.int64), tf.constant(vals, dtype=tf.float64), tf.constant([2, 6], dtype=tf.int64))
    got = run(tf.sparse_tensor_dense_matmul(sp, tf.constant(feats, dtype=tf.float64)))
    assert np.allclose(got, [feats[:2].mean(0), feats[2:].mean(0)])


def test_while_loop_traces_its_body_at_construction():
    # the reference rebinds the name its loop body closes over right after the call (feature_map.py:31-34);
    # TensorFlow has traced the body by then, and so must the shim
    x = tf.constant([[1., 2.], [3., 4.], [5., 6.]], dtype=tf.float64)
    n = tf.shape(x)[0]
    start = tf.expand_dims(x[0], 0)
    body = lambda i, acc: (i + 1, tf.concat([acc, tf.expand_dims(x[i], 0) * 10.0], 0))
    i, x = tf.while_loop(lambda i, acc: tf.less(i, n), body, (tf.constant(1), start))
    assert np.array_equal(run(x), [[1., 2.], [30., 40.], [50., 60.]])
    assert int(run(i)) == 3


def test_map_fn_and_tile_reshape_repeater():
    sizes = tf.constant([[2.], [3.]], dtype=tf.float64)
    emb = tf.constant([[1., 2., 3.], [4., 5., 6.]], dtype=tf.float64)
    size = tf.cast(sizes[1], tf.int32)                     # gp_utils.repeater
    rep = tf.reshape(tf.tile(emb[1], size), [tf.squeeze(size), tf.shape(emb[1])[0]])
    assert np.array_equal(run(rep), np.tile([4., 5., 6.], (3, 1)))
    sq = tf.map_fn(fn=lambda k: tf.reduce_sum(emb[k]) * 2.0, elems=tf.range_(2), dtype=tf.float64)
    assert np.array_equal(run(sq), [12., 30.])


def test_adam_step_is_tensorflows_update():
    tf.reset_default_graph()
    v = tf.Variable(tf.constant([1.0, -2.0], dtype=tf.float64), dtype=tf.float64)
    loss = tf.reduce_sum(tf.square(v) * tf.constant([1.0, 3.0], dtype=tf.float64))
    opt = tf.AdamOptimizer(0.1)
    step = opt.minimize(loss)
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    _, l0 = sess.run([step, loss])
    assert l0 == pytest.approx(1.0 + 12.0)                 # the fetched loss is the pre-update loss
    g = np.array([2.0, -12.0])
    m, s = 0.1 * g, 0.001 * g * g
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    want = np.array([1.0, -2.0]) - lr_t * m / (np.sqrt(s) + 1e-8)
    assert np.allclose(v.value.detach().numpy(), want, rtol=1e-14)
    gs, gn = tf.clip_by_global_norm(tf.gradients(loss, [v]), 5.0)
    clipped = sess.run(gs[0])
    w = v.value.detach().numpy()
    raw = np.array([2.0 * w[0], 6.0 * w[1]])
    assert np.allclose(clipped, raw * 5.0 / max(np.linalg.norm(raw), 5.0))
    tf.reset_default_graph()


@pytest.mark.skipif(not os.path.isdir("/root/reference/distBO"), reason="the reference tree exists only in the build container")
def test_committed_golden_files_regenerate_from_the_reference_sources(tmp_path):
    """Runs the reference's sources over the shim again (two fit cases, one BO trajectory) and compares with the
    committed files: the fixtures are the reference's numbers, not hand-edited ones."""
    code = (
        "import sys, os; sys.argv = ['x', '/root/reference', '--bo-only']\n"
        "sys.path.insert(0, %r)\n"
        "import make_reference_golden as M\n"
        "M.OUT = %r\n"
        "for n in ('params', 'dist_joint_clip', 'blr_manual'): M.run_case(n)\n"
        "M.run_bo_case('bo_dist_none')\n" % (os.path.join(ROOT, "tools"), str(tmp_path)))
    # '--bo-only' keeps the module's __main__ block from running everything; the cases are called explicitly
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    for name in ("ref_params", "ref_dist_joint_clip", "ref_blr_manual", "refbo_dist_none"):
        new = np.load(os.path.join(str(tmp_path), name + ".npz"))
        old = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        assert sorted(new.files) == sorted(old.files)
        for k in old.files:
            if k == "meta":
                assert str(new[k]) == str(old[k])
            else:
                assert np.allclose(new[k], old[k], rtol=1e-12, atol=1e-14, equal_nan=True), (name, k)
