"""Camera-gradient formulas of gsr_backward_camera (csrc/preprocess_bwd.cu, CAM variant), checked on the CPU against
PyTorch autograd in float64.

The per-Gaussian map  (viewmatrix, projmatrix, campos) -> (mean2D in NDC, conic, SH colour)  is smooth, so for FIXED
upstream gradients (dL/dmean2D, dL/dconic, dL/dcolour: what the blending backward hands to the preprocess backward) the
camera gradient is the plain chain rule. `camera_grads_formulas` restates the kernel's formulas in numpy, operation for
operation; `proxy_loss` restates the FORWARD map (cuda_rasterizer/forward.cu:74-113,196-200 and :20-71) in torch, and
autograd differentiates it. Agreement to 1e-9 pins the formulas (index conventions of the transposed matrices, the
W-in-T = W*J path, the p_hom.w quotient rule, the sign of the campos term); the GPU test then only has to show that the
CUDA transcription matches finite differences of the full pipeline within their (discontinuity-limited) accuracy."""
import numpy as np
import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199


def proxy_loss(view, proj, campos, means, cov3d, sh, g_mean2d, g_conic, g_col, hx, hy, limx, limy):
    """sum_i  g_mean2d . p_proj  +  g_conic . (A, B, C)  +  g_col . rgb   (degree-1 SH keeps the direction dependence)."""
    v, pm = view.reshape(-1), proj.reshape(-1)
    x, y, z = means[:, 0], means[:, 1], means[:, 2]
    tx = v[0] * x + v[4] * y + v[8] * z + v[12]
    ty = v[1] * x + v[5] * y + v[9] * z + v[13]
    tz = v[2] * x + v[6] * y + v[10] * z + v[14]
    hxw = pm[0] * x + pm[4] * y + pm[8] * z + pm[12]
    hyw = pm[1] * x + pm[5] * y + pm[9] * z + pm[13]
    hw = pm[3] * x + pm[7] * y + pm[11] * z + pm[15]
    m_w = 1.0 / (hw + 1e-7)
    px, py = hxw * m_w, hyw * m_w
    # no Gaussian of the test is clamped (|t.x/t.z| < limx): the clamp is the identity, as in the kernel's x_grad_mul = 1
    J00, J02, J11, J12 = hx / tz, -(hx * tx) / (tz * tz), hy / tz, -(hy * ty) / (tz * tz)
    W0 = torch.stack([v[0], v[4], v[8]]); W1 = torch.stack([v[1], v[5], v[9]]); W2 = torch.stack([v[2], v[6], v[10]])
    T0 = W0[None, :] * J00[:, None] + W2[None, :] * J02[:, None]      # T[0][r]
    T1 = W1[None, :] * J11[:, None] + W2[None, :] * J12[:, None]      # T[1][r]
    V = torch.zeros(means.shape[0], 3, 3, dtype=means.dtype)
    V[:, 0, 0], V[:, 0, 1], V[:, 0, 2] = cov3d[:, 0], cov3d[:, 1], cov3d[:, 2]
    V[:, 1, 0], V[:, 1, 1], V[:, 1, 2] = cov3d[:, 1], cov3d[:, 3], cov3d[:, 4]
    V[:, 2, 0], V[:, 2, 1], V[:, 2, 2] = cov3d[:, 2], cov3d[:, 4], cov3d[:, 5]
    a = torch.einsum("pi,pij,pj->p", T0, V, T0) + 0.3
    b = torch.einsum("pi,pij,pj->p", T0, V, T1)
    c = torch.einsum("pi,pij,pj->p", T1, V, T1) + 0.3
    det = a * c - b * b
    A, B, Cc = c / det, -b / det, a / det
    d = means - campos[None, :]
    d = d / d.norm(dim=1, keepdim=True)
    rgb = C0 * sh[:, 0] - C1 * d[:, 1:2] * sh[:, 1] + C1 * d[:, 2:3] * sh[:, 2] - C1 * d[:, 0:1] * sh[:, 3] + 0.5
    return (g_mean2d[:, 0] * px + g_mean2d[:, 1] * py).sum() + (g_conic[:, 0] * A + g_conic[:, 1] * B + g_conic[:, 2] * Cc).sum() \
        + (g_col * rgb).sum()


def camera_grads_formulas(view, proj, campos, means, cov3d, sh, g_mean2d, g_conic, g_col, hx, hy):
    """numpy transcription of the CAM blocks of preprocess_bwd_kernel (same variable names)."""
    vm, pr = view.reshape(-1), proj.reshape(-1)
    dview, dproj, dcam = np.zeros(16), np.zeros(16), np.zeros(3)
    for i in range(means.shape[0]):
        mean = means[i]
        t = np.array([vm[0] * mean[0] + vm[4] * mean[1] + vm[8] * mean[2] + vm[12],
                      vm[1] * mean[0] + vm[5] * mean[1] + vm[9] * mean[2] + vm[13],
                      vm[2] * mean[0] + vm[6] * mean[1] + vm[10] * mean[2] + vm[14]])
        J00, J02, J11, J12 = hx / t[2], -(hx * t[0]) / t[2] ** 2, hy / t[2], -(hy * t[1]) / t[2] ** 2
        W = np.array([[vm[0], vm[4], vm[8]], [vm[1], vm[5], vm[9]], [vm[2], vm[6], vm[10]]])   # W[c][r]
        T0 = W[0] * J00 + W[2] * J02
        T1 = W[1] * J11 + W[2] * J12
        c3 = cov3d[i]
        Vrk = np.array([[c3[0], c3[1], c3[2]], [c3[1], c3[3], c3[4]], [c3[2], c3[4], c3[5]]])
        ca, cb, cc = T0 @ Vrk @ T0 + 0.3, T0 @ Vrk @ T1, T1 @ Vrk @ T1 + 0.3
        gca, gcb, gcc = g_conic[i]
        denom = ca * cc - cb * cb
        d2 = 1.0 / (denom * denom)     # (the kernel adds 1e-7 to denom^2; immaterial here)
        dL_da = d2 * (-cc * cc * gca + 2 * cb * cc * gcb + (denom - ca * cc) * gcc)
        dL_dc = d2 * (-ca * ca * gcc + 2 * ca * cb * gcb + (denom - ca * cc) * gca)
        dL_db = d2 * 2 * (cb * cc * gca - (denom + 2 * cb * cb) * gcb + ca * cb * gcc)
        u0, u1 = Vrk @ T0, Vrk @ T1
        dT0 = 2 * u0 * dL_da + u1 * dL_db
        dT1 = 2 * u1 * dL_dc + u0 * dL_db
        dJ00, dJ02 = W[0] @ dT0, W[2] @ dT0
        dJ11, dJ12 = W[1] @ dT1, W[2] @ dT1
        tz = 1.0 / t[2]
        dt = np.array([-hx * tz * tz * dJ02, -hy * tz * tz * dJ12,
                       -hx * tz * tz * dJ00 - hy * tz * tz * dJ11 + 2 * hx * t[0] * tz ** 3 * dJ02 + 2 * hy * t[1] * tz ** 3 * dJ12])
        mj = np.array([mean[0], mean[1], mean[2], 1.0])
        for k in range(3):
            for j in range(4):
                dview[k + 4 * j] += dt[k] * mj[j]
        for r in range(3):
            dview[0 + 4 * r] += dT0[r] * J00
            dview[1 + 4 * r] += dT1[r] * J11
            dview[2 + 4 * r] += dT0[r] * J02 + dT1[r] * J12
        hw = pr[3] * mean[0] + pr[7] * mean[1] + pr[11] * mean[2] + pr[15]
        m_w = 1.0 / (hw + 1e-7)
        mul1 = (pr[0] * mean[0] + pr[4] * mean[1] + pr[8] * mean[2] + pr[12]) * m_w * m_w
        mul2 = (pr[1] * mean[0] + pr[5] * mean[1] + pr[9] * mean[2] + pr[13]) * m_w * m_w
        g2x, g2y = g_mean2d[i]
        dh = [g2x * m_w, g2y * m_w, -(g2x * mul1 + g2y * mul2)]
        for r, row in enumerate((0, 1, 3)):
            for j in range(4):
                dproj[row + 4 * j] += dh[r] * mj[j]
        do = mean - campos
        s2 = do @ do
        n = do / np.sqrt(s2)
        dRGB = g_col[i]
        ddir = np.array([-C1 * (sh[i, 3] @ dRGB), -C1 * (sh[i, 1] @ dRGB), C1 * (sh[i, 2] @ dRGB)])   # d rgb / d(unit dir)
        inv = 1.0 / np.sqrt(s2 ** 3)
        dmean_dir = np.array([((s2 - do[0] * do[0]) * ddir[0] - do[1] * do[0] * ddir[1] - do[2] * do[0] * ddir[2]) * inv,
                              (-do[0] * do[1] * ddir[0] + (s2 - do[1] * do[1]) * ddir[1] - do[2] * do[1] * ddir[2]) * inv,
                              (-do[0] * do[2] * ddir[0] - do[1] * do[2] * ddir[1] + (s2 - do[2] * do[2]) * ddir[2]) * inv])
        dcam -= dmean_dir
    return dview.reshape(4, 4), dproj.reshape(4, 4), dcam


def test_camera_gradient_formulas_equal_autograd_of_the_forward_map():
    from gaussianeditor_b200 import synth
    rng = np.random.default_rng(0)
    cam = synth.ring_cameras(8, 4.5, 15.0, 96, 64, 61.0)[3]
    P = 40
    means = rng.normal(0, 0.6, (P, 3))
    Lm = rng.normal(0, 0.05, (P, 3, 3))
    S = np.einsum("pij,pkj->pik", Lm, Lm) + 1e-4 * np.eye(3)
    cov3d = np.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1)
    sh = rng.normal(0, 0.5, (P, 4, 3))
    g_m2, g_con, g_col = rng.normal(size=(P, 2)), rng.normal(size=(P, 3)), rng.normal(size=(P, 3))
    hx, hy = cam.image_width / (2 * cam.tanfovx), cam.image_height / (2 * cam.tanfovy)
    view, proj, cpos = cam.viewmatrix.astype(np.float64), cam.projmatrix.astype(np.float64), cam.campos.astype(np.float64)
    tv = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (view, proj, cpos)]
    t = lambda a: torch.tensor(a, dtype=torch.float64)
    loss = proxy_loss(tv[0], tv[1], tv[2], t(means), t(cov3d), t(sh), t(g_m2), t(g_con), t(g_col), hx, hy,
                      1.3 * cam.tanfovx, 1.3 * cam.tanfovy)
    loss.backward()
    # The reference's blending backward hands over HALF of dL/d(conic b): it accumulates -0.5 * dx * dy * dL/dG
    # (backward.cu:550) although d(power)/db = -dx*dy, and computeCov2DCUDA compensates with the factor 2 in dL_db
    # (backward.cu:212). The kernel keeps that convention, so the formulas get g_conic_b / 2.
    g_con_ref = g_con * np.array([1.0, 0.5, 1.0])
    dv, dp, dc = camera_grads_formulas(view, proj, cpos, means, cov3d, sh, g_m2, g_con_ref, g_col, hx, hy)
    for got, want in ((dv, tv[0].grad.numpy()), (dp, tv[1].grad.numpy()), (dc, tv[2].grad.numpy())):
        assert np.linalg.norm(want) > 0
        assert np.linalg.norm(got - want) <= 1e-9 * np.linalg.norm(want), (got, want)
    assert np.all(dv[:, 3] == 0) and np.all(dp[:, 2] == 0)
