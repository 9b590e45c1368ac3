"""GPU leg of tests/golden/validation_edge.npz (oracle/gen_golden_edge.py): the drop-in FaceToFaceValidation through the C ABI
against outputs of the UNMODIFIED reference on ragged folds (N % k != 0), test folds without a same-identity pair and a
two-class set under metric 1.

The fixture's seeds keep every pair distance more than 3e-5 away from every grid threshold and from every chosen
threshold, so with distances inside the 1e-5 tolerance the integer counts -- and therefore the chosen thresholds -- are
the reference's EXACTLY; what is left is the fp64 summation order of the class-pair weights (~1e-15).
(Written after round 1's GPU budget was spent; the host-logic leg over the emulated library runs in the CPU suite:
tests/test_host_logic.py::test_validation_edge_cases_host_math_golden.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_validation_edge_cases_golden(golden_dir):
    from facenet_b200 import statistics as fst
    g = np.load(golden_dir / 'validation_edge.npz')
    for name in (str(c) for c in g['cases']):
        x, labels = g[name + '_embeddings'], g[name + '_labels']

        class Cfg:
            metric, nrof_folds, far_target = int(g[name + '_cfg'][0]), int(g[name + '_cfg'][1]), float(g[name + '_cfg'][2])

        v = fst.FaceToFaceValidation(x, labels, Cfg)
        assert sorted(v.dict.keys()) == [str(k) for k in g[name + '_criteria']]
        for r, tag in zip(v.reports, ('acc', 'far')):
            dct = r.dict
            keys = [str(k) for k in g['%s_%s_keys' % (name, tag)]]
            assert sorted(dct.keys()) == keys
            got_thr = np.array([float(m.threshold[0]) for m in r.conf_matrix_test])
            if tag == 'acc':
                np.testing.assert_array_equal(got_thr, g[name + '_acc_thr'], err_msg=name)          # grid points
            else:
                np.testing.assert_allclose(got_thr, g[name + '_far_thr'], rtol=0, atol=1e-9, err_msg=name)
            got_test = np.array([[m.tp[0], m.tn[0], m.fp[0], m.fn[0]] for m in r.conf_matrix_test])
            np.testing.assert_allclose(got_test, g['%s_%s_test' % (name, tag)], rtol=0, atol=1e-12, err_msg=name)
            np.testing.assert_allclose([float(dct[k]) for k in keys], g['%s_%s_vals' % (name, tag)], rtol=0, atol=1e-6,
                                       err_msg='%s %s' % (name, tag))
        np.testing.assert_allclose(np.array([m.tp for m in v.reports[0].conf_matrix_train]), g[name + '_train_tp'], rtol=0, atol=1e-12)
        np.testing.assert_allclose(np.array([m.fp for m in v.reports[0].conf_matrix_train]), g[name + '_train_fp'], rtol=0, atol=1e-12)
        assert repr(v).split('elapsed_time')[0].splitlines()[:3] == str(g[name + '_repr']).splitlines()[:3]
