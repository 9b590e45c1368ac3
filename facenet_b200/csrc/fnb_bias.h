// facenet_b200 -- calibration of the tensor core's accumulation bias and of the per-mode error model (host side).
//
// MEASURED on B200 (profiles/r02a_bias_*.log, scripts/probe_bias.py: 33.5 M pairs of 8192 x D rows, classes of graded tightness,
// float64 reference): tcgen05.mma truncates towards zero when it aligns the products and adds them into the fp32 accumulator,
// so a raw similarity is smaller in magnitude than the exact one, s_raw = s (1 - beta(|s|)).  beta is 1.7e-6 .. 3.3e-6 for the
// 96 accumulation steps of fp16x3 at D = 512 (max |dd| 9.2e-6 -- the "fp32-equivalent" split would sit at the edge of the
// 1e-5 contract, and beyond it at D = 1024: 1.75e-5), 0.5e-6 .. 2.3e-6 for fp16f8 (cross terms first: 32 lossy steps), and
// grows ~ D^1.09 (D = 128 / 256 / 512 / 1024: 0.65 / 1.43 / 3.05 / 6.35e-6 for fp16x3).  The spread around the mean is small
// (fp16x3: 0.05 .. 0.3e-6), so correcting the mean recovers fp32-like accuracy.  beta also depends on the rows (all-positive
// rows +30 %, rows with 7/8 zeros -50 %): the tables are for dense zero-mean coordinates, what an embedding network with
// l2_normalize produces; the residual on other data stays below ~1e-6 |s|.
#pragma once

namespace fnb {

// beta(|s|) at |s| = 0, 0.025, ..., 1 for D = 512 (bin means of the probe, [1 2 1] / 4 smoothed)
static const float kBiasBetaX3[41] = {
    1.668e-6f, 1.664e-6f, 1.663e-6f, 1.663e-6f, 1.660e-6f, 1.698e-6f, 1.765e-6f, 1.795e-6f, 1.777e-6f, 1.765e-6f, 1.820e-6f,
    1.918e-6f, 1.996e-6f, 2.047e-6f, 2.076e-6f, 2.081e-6f, 2.075e-6f, 2.064e-6f, 2.048e-6f, 2.061e-6f, 2.138e-6f, 2.247e-6f,
    2.346e-6f, 2.432e-6f, 2.490e-6f, 2.519e-6f, 2.556e-6f, 2.620e-6f, 2.684e-6f, 2.734e-6f, 2.788e-6f, 2.830e-6f, 2.844e-6f,
    2.869e-6f, 2.913e-6f, 2.952e-6f, 2.992e-6f, 3.049e-6f, 3.146e-6f, 3.261e-6f, 3.366e-6f};
static const float kBiasBetaF8[41] = {
    0.529e-6f, 0.529e-6f, 0.537e-6f, 0.552e-6f, 0.564e-6f, 0.596e-6f, 0.646e-6f, 0.676e-6f, 0.706e-6f, 0.736e-6f, 0.763e-6f,
    0.809e-6f, 0.852e-6f, 0.867e-6f, 0.869e-6f, 0.869e-6f, 0.873e-6f, 0.897e-6f, 0.938e-6f, 0.974e-6f, 1.010e-6f, 1.059e-6f,
    1.129e-6f, 1.215e-6f, 1.294e-6f, 1.348e-6f, 1.397e-6f, 1.462e-6f, 1.537e-6f, 1.607e-6f, 1.657e-6f, 1.670e-6f, 1.676e-6f,
    1.723e-6f, 1.801e-6f, 1.869e-6f, 1.927e-6f, 1.989e-6f, 2.081e-6f, 2.199e-6f, 2.318e-6f};

}  // namespace fnb
