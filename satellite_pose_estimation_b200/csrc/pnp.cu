// Set post-processing + pose solve, one warp per image, fp64.
//
// Replaces, per batch, the reference's D2H copy + Python loop:
//   PostProcess.forward            softmax, de-normalise keypoints to image pixels     RV/models/detr_speed.py:264-293
//   SimplePoseSolver.__call__      argmax label / score per query, drop background,    RV/utils/speed_eval.py:164-242
//                                  keep the best-scoring query per label,
//                                  cv2.solvePnPRansac(P3P, reproj) -> inlier set,
//                                  cv2.solvePnPGeneric(ITERATIVE) LM refinement on the inliers,
//                                  Rodrigues -> quaternion (w,x,y,z)
//   SimplePoseSolverSigma          same with sigma-weighted Huber LM in normalised     SA/utils/speed_eval.py:269-420
//                                  image coordinates, w = 1/(sqrt(sigma)+1e-6) / sum
//   Multi_Mean_PoseSolver          ensemble of N checkpoints: every foreground query   RV/utils/speed_eval.py:42-140
//                                  of every model is pooled per label, mean -> 3-sigma
//                                  distance filter -> mean, then the same PnP chain
//
// cv2's RANSAC draws random 4-point samples; the warp instead evaluates EVERY 3-point minimal sample (<= 165 for 11
// keypoints; Grunert's P3P, closed-form quartic) against all correspondences and keeps the hypothesis with the most
// inliers (ties: lowest inlier error).  That is the consensus set RANSAC converges to, obtained deterministically.
// The refinement is Levenberg-Marquardt on the 6-DoF pose (left-multiplicative so(3) update), lanes = points,
// warp-shuffle reductions for the normal equations, every lane solving the 6x6 system redundantly.
// Failures reproduce the reference's observable contract: < 4 correspondences / no valid pose -> zero pose + status.
#include "spe_internal.h"
#include "profile.h"
#include "spe_ptx.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace spe {

namespace {

__constant__ double c_world[11 * 3] = {
    0.30531443180639595,  -0.5789365328147589, 0.25084064329219374,   0.5447777012552748,   0.4896098588217239,
    0.2527042917812688,   -0.5428973667440873, 0.4888589385025832,    0.25350052140860274,  0.3666281919575766,
    -0.3823462337812798,  0.3221231197241823,  0.3648084120091035,    0.38159211256229386,  0.3198573872530155,
    -0.36705288820278714, 0.38095878832554714, 0.32031160558604727,   -0.3671484046314764,  -0.3815359857639992,
    0.3209066585512606,   0.3673520558953431,  -0.2620043692501464,   0.001723572896525486, 0.36711999898725295,
    0.30142490961836477,  -0.00013418389188803165, -0.36787140119087025, 0.3015820378676121, 0.0012482861217676527,
    -0.3679806481789124,  -0.2621021059553393, 0.0006999278181541126};  // RV/all_result.json "pt" fields

// Camera: RV/utils/utils.py:30-46
constexpr double kFx = 0.0176 / 5.86e-6;
constexpr double kFy = 0.0176 / 5.86e-6;
constexpr double kCx = 960.0;
constexpr double kCy = 600.0;

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ---- closed-form real roots -----------------------------------------------------------------------------------
__device__ double cubic_one_real(double a2, double a1, double a0) {  // x^3 + a2 x^2 + a1 x + a0
  const double q = (3.0 * a1 - a2 * a2) / 9.0;
  const double r = (9.0 * a2 * a1 - 27.0 * a0 - 2.0 * a2 * a2 * a2) / 54.0;
  const double d = q * q * q + r * r;
  if (d >= 0.0) {
    const double sd = sqrt(d);
    return cbrt(r + sd) + cbrt(r - sd) - a2 / 3.0;
  }
  double c = r / sqrt(-q * q * q);
  c = fmin(1.0, fmax(-1.0, c));
  return 2.0 * sqrt(-q) * cos(acos(c) / 3.0) - a2 / 3.0;
}

__device__ int quartic_real_roots(double c4, double c3, double c2, double c1, double c0, double (&x)[4]) {
  const double big = fmax(fmax(fabs(c3), fabs(c2)), fmax(fabs(c1), fmax(fabs(c0), 1e-300)));
  if (fabs(c4) < 1e-14 * big) return 0;
  const double a = c3 / c4, b = c2 / c4, c = c1 / c4, d = c0 / c4;
  const double a2 = a * a;
  const double p = b - 3.0 * a2 / 8.0;
  const double q = c - a * b / 2.0 + a2 * a / 8.0;
  const double r = d - a * c / 4.0 + a2 * b / 16.0 - 3.0 * a2 * a2 / 256.0;
  int n = 0;
  if (fabs(q) < 1e-14) {
    const double disc = p * p - 4.0 * r;
    if (disc >= 0.0) {
      const double sd = sqrt(disc);
      const double z0 = (-p + sd) / 2.0, z1 = (-p - sd) / 2.0;
      if (z0 >= 0.0) { const double s = sqrt(z0); x[n++] = s; x[n++] = -s; }
      if (z1 >= 0.0) { const double s = sqrt(z1); x[n++] = s; x[n++] = -s; }
    }
  } else {
    const double m = cubic_one_real(p, p * p / 4.0 - r, -q * q / 8.0);
    if (!(m > 0.0)) return 0;
    const double s = sqrt(2.0 * m);
#pragma unroll
    for (int sg = 0; sg < 2; ++sg) {
      const double sign = sg == 0 ? 1.0 : -1.0;
      const double bb = sign * s;
      const double cc = p / 2.0 + m - sign * q / (2.0 * s);
      const double disc = bb * bb - 4.0 * cc;
      if (disc >= 0.0) {
        const double sd = sqrt(disc);
        x[n++] = (-bb + sd) / 2.0;
        x[n++] = (-bb - sd) / 2.0;
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    double v = x[i] - a / 4.0;
#pragma unroll
    for (int it = 0; it < 2; ++it) {  // Newton polish on the original polynomial
      const double f = (((c4 * v + c3) * v + c2) * v + c1) * v + c0;
      const double df = ((4.0 * c4 * v + 3.0 * c3) * v + 2.0 * c2) * v + c1;
      if (df != 0.0) v -= f / df;
    }
    x[i] = v;
  }
  return n;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// orthonormal frame (columns e1,e2,e3) of the triangle X0,X1,X2; returns false if degenerate
__device__ bool tri_frame(const double* X0, const double* X1, const double* X2, double (&F)[9]) {
  double e1[3] = {X1[0] - X0[0], X1[1] - X0[1], X1[2] - X0[2]};
  double d2[3] = {X2[0] - X0[0], X2[1] - X0[1], X2[2] - X0[2]};
  double e3[3];
  cross3(e1, d2, e3);
  const double n1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
  const double n3 = sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
  if (n1 < 1e-12 || n3 < 1e-12) return false;
#pragma unroll
  for (int i = 0; i < 3; ++i) { e1[i] /= n1; e3[i] /= n3; }
  double e2[3];
  cross3(e3, e1, e2);
#pragma unroll
  for (int i = 0; i < 3; ++i) { F[i * 3 + 0] = e1[i]; F[i * 3 + 1] = e2[i]; F[i * 3 + 2] = e3[i]; }
  return true;
}

struct Hyp {
  int cnt;
  double err;
  unsigned mask;
  double R[9];
  double t[3];
};

// squared pixel reprojection errors of all n correspondences under (R,t) -> inlier count / error sum / mask
__device__ void score_pose(const double (&R)[9], const double (&t)[3], int n, const int* lab, const double* uv,
                           double thr2, int& cnt, double& err, unsigned& mask) {
  cnt = 0; err = 0.0; mask = 0u;
  for (int i = 0; i < n; ++i) {
    const double* X = c_world + lab[i] * 3;
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    if (!(z > 0.0)) continue;
    const double du = kFx * x / z + kCx - uv[2 * i], dv = kFy * y / z + kCy - uv[2 * i + 1];
    const double e = du * du + dv * dv;
    if (e <= thr2) { ++cnt; err += e; mask |= 1u << i; }
  }
}

// Grunert's three-point pose: all admissible (R,t) for correspondences (i0,i1,i2); each one is scored immediately
__device__ void p3p_consensus(int i0, int i1, int i2, int n, const int* lab, const double* uv, const double* bear,
                              double thr2, Hyp& best) {
  const double* P0 = c_world + lab[i0] * 3;
  const double* P1 = c_world + lab[i1] * 3;
  const double* P2 = c_world + lab[i2] * 3;
  const double* f0 = bear + 3 * i0;
  const double* f1 = bear + 3 * i1;
  const double* f2 = bear + 3 * i2;
  auto d2 = [](const double* a, const double* b) {
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return x * x + y * y + z * z;
  };
  auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  const double a2 = d2(P1, P2), b2 = d2(P0, P2), c2 = d2(P0, P1);
  if (b2 < 1e-18) return;
  const double ca = dot(f1, f2), cb = dot(f0, f2), cg = dot(f0, f1);
  const double K1 = (a2 - c2) / b2, K2 = (a2 + c2) / b2;
  const double A4 = (K1 - 1.0) * (K1 - 1.0) - 4.0 * c2 / b2 * ca * ca;
  const double A3 = 4.0 * (K1 * (1.0 - K1) * cb - (1.0 - K2) * ca * cg + 2.0 * c2 / b2 * ca * ca * cb);
  const double A2 = 2.0 * (K1 * K1 - 1.0 + 2.0 * K1 * K1 * cb * cb + 2.0 * ((b2 - c2) / b2) * ca * ca -
                           4.0 * K2 * ca * cb * cg + 2.0 * ((b2 - a2) / b2) * cg * cg);
  const double A1 = 4.0 * (-K1 * (1.0 + K1) * cb + 2.0 * a2 / b2 * cg * cg * cb - (1.0 - K2) * ca * cg);
  const double A0 = (1.0 + K1) * (1.0 + K1) - 4.0 * a2 / b2 * cg * cg;
  double roots[4];
  const int nr = quartic_real_roots(A4, A3, A2, A1, A0, roots);
  double Fw[9];
  if (!tri_frame(P0, P1, P2, Fw)) return;
  for (int ri = 0; ri < nr; ++ri) {
    const double v = roots[ri];
    if (!(v > 0.0)) continue;
    const double den = 2.0 * (cg - v * ca);
    if (fabs(den) < 1e-12) continue;
    const double u = ((K1 - 1.0) * v * v - 2.0 * K1 * cb * v + 1.0 + K1) / den;
    if (!(u > 0.0)) continue;
    const double dd = 1.0 + v * v - 2.0 * v * cb;
    if (!(dd > 0.0)) continue;
    const double s1 = sqrt(b2 / dd), s2 = u * s1, s3 = v * s1;
    const double C0[3] = {s1 * f0[0], s1 * f0[1], s1 * f0[2]};
    const double C1[3] = {s2 * f1[0], s2 * f1[1], s2 * f1[2]};
    const double C2[3] = {s3 * f2[0], s3 * f2[1], s3 * f2[2]};
    double Fc[9];
    if (!tri_frame(C0, C1, C2, Fc)) continue;
    double R[9], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        R[i * 3 + j] = Fc[i * 3 + 0] * Fw[j * 3 + 0] + Fc[i * 3 + 1] * Fw[j * 3 + 1] + Fc[i * 3 + 2] * Fw[j * 3 + 2];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = C0[i] - (R[i * 3] * P0[0] + R[i * 3 + 1] * P0[1] + R[i * 3 + 2] * P0[2]);
    int cnt; double err; unsigned mask;
    score_pose(R, t, n, lab, uv, thr2, cnt, err, mask);
    if (cnt > best.cnt || (cnt == best.cnt && cnt > 0 && err < best.err)) {
      best.cnt = cnt; best.err = err; best.mask = mask;
#pragma unroll
      for (int i = 0; i < 9; ++i) best.R[i] = R[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) best.t[i] = t[i];
    }
  }
}

// R <- exp([w]x) R
__device__ void rot_update(const double* w, const double (&R)[9], double (&Rn)[9]) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double A, Bc;
  if (th < 1e-8) { A = 1.0 - th2 / 6.0; Bc = 0.5 - th2 / 24.0; }
  else { A = sin(th) / th; Bc = (1.0 - cos(th)) / th2; }
  // E = I + A K + Bc K^2
  const double K[9] = {0.0, -w[2], w[1], w[2], 0.0, -w[0], -w[1], w[0], 0.0};
  double E[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double k2 = K[i * 3 + 0] * K[0 * 3 + j] + K[i * 3 + 1] * K[1 * 3 + j] + K[i * 3 + 2] * K[2 * 3 + j];
      E[i * 3 + j] = (i == j ? 1.0 : 0.0) + A * K[i * 3 + j] + Bc * k2;
    }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Rn[i * 3 + j] = E[i * 3 + 0] * R[0 * 3 + j] + E[i * 3 + 1] * R[1 * 3 + j] + E[i * 3 + 2] * R[2 * 3 + j];
}

// solve (A + lam diag(A)) d = -g for symmetric 6x6 A (upper triangle packed row-major, 21 values); false if not PD
__device__ bool solve6(const double* Ap, const double* g, double lam, double* d) {
  double L[6][6];
  int idx = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) { L[j][i] = Ap[idx]; L[i][j] = Ap[idx]; ++idx; }
  for (int i = 0; i < 6; ++i) L[i][i] *= (1.0 + lam);
  for (int j = 0; j < 6; ++j) {
    double s = L[j][j];
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    if (!(s > 1e-300)) return false;
    const double ljj = sqrt(s);
    L[j][j] = ljj;
    for (int i = j + 1; i < 6; ++i) {
      double v = L[i][j];
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v / ljj;
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double v = -g[i];
    for (int k = 0; k < i; ++k) v -= L[i][k] * y[k];
    y[i] = v / L[i][i];
  }
  for (int i = 5; i >= 0; --i) {
    double v = y[i];
    for (int k = i + 1; k < 6; ++k) v -= L[k][i] * d[k];
    d[i] = v / L[i][i];
  }
  return true;
}

// residual model shared by the plain (pixel) and sigma-weighted (normalised coordinates) solvers:
//   r_u = wu * (au * x/z + cu - mu),  r_v = wv * (av * y/z + cv - mv),  robust weight rho'(|r|^2) (Huber)
struct PointObs {
  double X[3];
  double mu, mv, wu, wv;
  bool active;
};

__device__ __forceinline__ double huber_rho(double s, double delta) {  // Ceres HuberLoss rho(s)
  return s <= delta * delta ? s : 2.0 * delta * sqrt(s) - delta * delta;
}

__device__ double pose_cost(const double (&R)[9], const double (&t)[3], const PointObs& o, double au, double cu,
                            double av, double cv, double huber, bool& zok) {
  double c = 0.0;
  bool ok = true;
  if (o.active) {
    const double x = R[0] * o.X[0] + R[1] * o.X[1] + R[2] * o.X[2] + t[0];
    const double y = R[3] * o.X[0] + R[4] * o.X[1] + R[5] * o.X[2] + t[1];
    const double z = R[6] * o.X[0] + R[7] * o.X[1] + R[8] * o.X[2] + t[2];
    ok = z > 0.0;
    const double ru = o.wu * (au * x / z + cu - o.mu), rv = o.wv * (av * y / z + cv - o.mv);
    const double s = ru * ru + rv * rv;
    c = huber > 0.0 ? huber_rho(s, huber) : s;
  }
  zok = __all_sync(FULL, ok);
  return wsum(c);
}

__device__ bool lm_refine(double (&R)[9], double (&t)[3], const PointObs& o, double au, double cu, double av,
                          double cv, double huber, int n, double* sJ /*[16][14]*/, double* sA /*[27]*/) {
  const int lane = threadIdx.x & 31;
  bool zok;
  double cost = pose_cost(R, t, o, au, cu, av, cv, huber, zok);
  if (!zok || !isfinite(cost)) return false;
  double lam = 1e-3;
  // entry e of the packed normal equations handled by lane e: e < 21 -> A(i,j) upper triangle, 21..26 -> g(e-21)
  int ei = 0, ej = 0;
  if (lane < 21) {
    int idx = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) { if (idx == lane) { ei = i; ej = j; } ++idx; }
  } else {
    ei = lane - 21;
  }
  for (int iter = 0; iter < 20; ++iter) {
    // lanes = points: residuals and Jacobian rows, pre-scaled by sqrt(rho') so that A = J^T J, g = J^T r
    if (lane < 16) {
      double Ju[6] = {0, 0, 0, 0, 0, 0}, Jv[6] = {0, 0, 0, 0, 0, 0}, ru = 0.0, rv = 0.0;
      if (o.active) {
        const double Y0 = R[0] * o.X[0] + R[1] * o.X[1] + R[2] * o.X[2];
        const double Y1 = R[3] * o.X[0] + R[4] * o.X[1] + R[5] * o.X[2];
        const double Y2 = R[6] * o.X[0] + R[7] * o.X[1] + R[8] * o.X[2];
        const double x = Y0 + t[0], y = Y1 + t[1], z = Y2 + t[2];
        const double iz = 1.0 / z;
        ru = o.wu * (au * x * iz + cu - o.mu);
        rv = o.wv * (av * y * iz + cv - o.mv);
        const double s = ru * ru + rv * rv;
        const double sw = (huber > 0.0 && s > huber * huber) ? sqrt(huber / sqrt(s)) : 1.0;  // sqrt(rho'(s))
        // d(point)/d(omega) = -[Y]x ; d(point)/dt = I ; du/d(point) = wu*au*[1/z, 0, -x/z^2]
        const double gu0 = sw * o.wu * au * iz, gu2 = -sw * o.wu * au * x * iz * iz;
        const double gv1 = sw * o.wv * av * iz, gv2 = -sw * o.wv * av * y * iz * iz;
        Ju[0] = gu2 * Y1;              Ju[1] = gu0 * Y2 - gu2 * Y0;   Ju[2] = -gu0 * Y1;
        Jv[0] = -gv1 * Y2 + gv2 * Y1;  Jv[1] = -gv2 * Y0;             Jv[2] = gv1 * Y0;
        Ju[3] = gu0; Ju[5] = gu2;
        Jv[4] = gv1; Jv[5] = gv2;
        ru *= sw; rv *= sw;
      }
      double* row = sJ + lane * 14;
#pragma unroll
      for (int i = 0; i < 6; ++i) { row[i] = Ju[i]; row[6 + i] = Jv[i]; }
      row[12] = ru; row[13] = rv;
    }
    __syncwarp();
    if (lane < 27) {
      double acc = 0.0;
      for (int pnt = 0; pnt < n; ++pnt) {
        const double* row = sJ + pnt * 14;
        if (lane < 21) acc += row[ei] * row[ej] + row[6 + ei] * row[6 + ej];
        else acc += row[ei] * row[12] + row[6 + ei] * row[13];
      }
      sA[lane] = acc;
    }
    __syncwarp();
    double Ap[21], g[6];
#pragma unroll
    for (int i = 0; i < 21; ++i) Ap[i] = sA[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = sA[21 + i];
    __syncwarp();
    bool accepted = false, converged = false;
    for (int tries = 0; tries < 8; ++tries) {
      double d[6];
      if (solve6(Ap, g, lam, d)) {
        double Rn[9], tn[3] = {t[0] + d[3], t[1] + d[4], t[2] + d[5]};
        rot_update(d, R, Rn);
        bool zk;
        const double cn = pose_cost(Rn, tn, o, au, cu, av, cv, huber, zk);
        if (zk && cn <= cost) {
          const double dn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
          converged = (cost - cn) <= 1e-13 * fmax(cost, 1e-300) || dn < 1e-11;
#pragma unroll
          for (int i = 0; i < 9; ++i) R[i] = Rn[i];
          t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
          cost = cn;
          lam = fmax(lam * 0.1, 1e-12);
          accepted = true;
          break;
        }
      }
      lam *= 10.0;
    }
    if (!accepted || converged) break;
  }
  return isfinite(cost);
}

__device__ void rot_to_quat(const double (&R)[9], double (&q)[4]) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) {
    const double s = sqrt(tr + 1.0) * 2.0;
    q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0;
    q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0;
    q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
  } else {
    const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0;
    q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
  }
  const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double sg = (q[0] < 0.0 ? -1.0 : 1.0) / nrm;
  q[0] *= sg; q[1] *= sg; q[2] *= sg; q[3] *= sg;
}

constexpr int kPnpThreads = 192;   // 165 triples of 11 correspondences in ONE round of the consensus loop (128: two)
constexpr int kPnpWarps = kPnpThreads / 32;
constexpr int kMaxPooled = 4096;     // ensemble: models x queries per image

__global__ void __launch_bounds__(kPnpThreads)
assign_pnp_kernel(const PnpDesc d) {
  pdl_wait();
  pdl_launch();
  const int img = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Q = d.Q;
  __shared__ int s_n;
  __shared__ int s_lab[11];
  __shared__ double s_uv[22];
  __shared__ double s_sig[22];
  __shared__ double s_bear[33];
  __shared__ int s_bcnt[kPnpWarps];
  __shared__ double s_berr[kPnpWarps];
  __shared__ unsigned s_bmask[kPnpWarps];
  __shared__ double s_bpose[kPnpWarps][12];
  __shared__ double s_J[16 * 14];
  __shared__ double s_A[27];
  __shared__ unsigned char s_elab[kMaxPooled];   // ensemble: label of every pooled prediction

  const float* lg = d.logits + static_cast<long long>(img) * Q * 12;
  const float* pt = d.points + static_cast<long long>(img) * Q * 2;
  // PostProcess multiplies the fp32 points by the box extent and adds the box origin, all in fp32 (a 0-dim int64 or
  // float64 box tensor is cast to the points' dtype first): the int boxes of the submission path, or the unrounded
  // float box of the eval path
  float bx1, by1, bw, bh;
  if (d.boxes_f) {
    bx1 = d.boxes_f[img * 4 + 0]; by1 = d.boxes_f[img * 4 + 1];
    bw = d.boxes_f[img * 4 + 2]; bh = d.boxes_f[img * 4 + 3];
  } else {
    const int ix1 = d.boxes[img * 4 + 0], iy1 = d.boxes[img * 4 + 1];
    bx1 = static_cast<float>(ix1); by1 = static_cast<float>(iy1);
    bw = static_cast<float>(d.boxes[img * 4 + 2] - ix1);
    bh = static_cast<float>(d.boxes[img * 4 + 3] - iy1);
  }

  const long long t_start = clock64();
  if (warp == 0 && d.num_models > 0) {
    // ---- ensemble: Multi_Mean_PoseSolver.__call__ / mean_and_filter (RV/utils/speed_eval.py:57-96)
    // pooled prediction e = model * Q + query, i.e. the order in which the reference appends to obj_pts_original
    const int NQ = d.num_models * Q;
    auto pixel = [&](int e, float& px, float& py) {
      const int m = e / Q, q = e - m * Q;
      const float* p2 = d.points + ((static_cast<long long>(m) * d.B + img) * Q + q) * 2;
      px = __fadd_rn(__fmul_rn(p2[0], bw), bx1);     // PostProcess, fp32, unfused
      py = __fadd_rn(__fmul_rn(p2[1], bh), by1);
    };
    for (int e = lane; e < NQ; e += 32) {
      const int m = e / Q, q = e - m * Q;
      const float* x = d.logits + ((static_cast<long long>(m) * d.B + img) * Q + q) * 12;
      float mx = -INFINITY;
      int am = 0;
#pragma unroll
      for (int c = 0; c < 12; ++c) {
        const float v = x[c];
        if (v > mx) { mx = v; am = c; }     // first maximum, like np.argmax
      }
      s_elab[e] = static_cast<unsigned char>(am);
    }
    __syncwarp();
    float mx = 0.f, my = 0.f;
    int cnt = 0, first = 0x7fffffff;
    if (lane < 11) {
      // np.mean over float32 rows: sequential fp32 sums in pooling order, one fp32 division
      float sx = 0.f, sy = 0.f;
      int n = 0;
      for (int e = 0; e < NQ; ++e) {
        if (s_elab[e] != lane) continue;
        float px, py;
        pixel(e, px, py);
        sx = __fadd_rn(sx, px); sy = __fadd_rn(sy, py);
        if (n == 0) first = e;
        ++n;
      }
      if (n > 0) {
        const float m0x = __fdiv_rn(sx, static_cast<float>(n)), m0y = __fdiv_rn(sy, static_cast<float>(n));
        if (n < 3) {
          mx = m0x; my = m0y; cnt = n;
        } else {
          // cdist(points, mean) in float64, np.std of the distances, keep distances < 3 std, mean again
          double sd = 0.0;
          for (int e = 0; e < NQ; ++e) {
            if (s_elab[e] != lane) continue;
            float px, py;
            pixel(e, px, py);
            const double dx = static_cast<double>(px) - static_cast<double>(m0x);
            const double dy = static_cast<double>(py) - static_cast<double>(m0y);
            sd += sqrt(dx * dx + dy * dy);
          }
          const double md = sd / n;
          double var = 0.0;
          for (int e = 0; e < NQ; ++e) {
            if (s_elab[e] != lane) continue;
            float px, py;
            pixel(e, px, py);
            const double dx = static_cast<double>(px) - static_cast<double>(m0x);
            const double dy = static_cast<double>(py) - static_cast<double>(m0y);
            const double dd = sqrt(dx * dx + dy * dy) - md;
            var += dd * dd;
          }
          const double thr = 3.0 * sqrt(var / n);
          float s2x = 0.f, s2y = 0.f;
          int k = 0;
          for (int e = 0; e < NQ; ++e) {
            if (s_elab[e] != lane) continue;
            float px, py;
            pixel(e, px, py);
            const double dx = static_cast<double>(px) - static_cast<double>(m0x);
            const double dy = static_cast<double>(py) - static_cast<double>(m0y);
            if (sqrt(dx * dx + dy * dy) < thr) { s2x = __fadd_rn(s2x, px); s2y = __fadd_rn(s2y, py); ++k; }
          }
          // k == 0 (all distances equal, e.g. coincident predictions): the reference averages an empty set -> NaN
          // image point, which its RANSAC can never count as an inlier; here the label is dropped instead
          if (k > 0) { mx = __fdiv_rn(s2x, static_cast<float>(k)); my = __fdiv_rn(s2y, static_cast<float>(k)); cnt = k; }
        }
      }
    }
    const bool present = lane < 11 && cnt > 0;
    // correspondence order = order of first appearance of each label in the pooled list (dict insertion order)
    int rank = 0;
    for (int l = 0; l < 11; ++l) {
      const int f2 = __shfl_sync(FULL, first, l);
      const int c2 = __shfl_sync(FULL, cnt, l);
      if (c2 > 0 && f2 < first) ++rank;
    }
    const unsigned pm = __ballot_sync(FULL, present);
    if (lane < 11) {
      d.assign[img * 11 + lane] = cnt;
      if (d.pooled_px) {
        d.pooled_px[(img * 11 + lane) * 2 + 0] = present ? mx : 0.f;
        d.pooled_px[(img * 11 + lane) * 2 + 1] = present ? my : 0.f;
      }
    }
    if (present) {
      s_lab[rank] = lane;
      s_uv[2 * rank] = static_cast<double>(mx);
      s_uv[2 * rank + 1] = static_cast<double>(my);
      const double bx = (static_cast<double>(mx) - kCx) / kFx, by = (static_cast<double>(my) - kCy) / kFy;
      const double inv = 1.0 / sqrt(bx * bx + by * by + 1.0);
      s_bear[3 * rank] = bx * inv; s_bear[3 * rank + 1] = by * inv; s_bear[3 * rank + 2] = inv;
      s_sig[2 * rank] = 1.0; s_sig[2 * rank + 1] = 1.0;
    }
    if (lane == 0) s_n = __popc(pm);
  } else if (warp == 0) {
    // ---- PostProcess + find_index, one query per lane per pass; per-label running best (score desc, query asc)
    float best_s[11];
    int best_q[11];
#pragma unroll
    for (int l = 0; l < 11; ++l) { best_s[l] = -1.f; best_q[l] = 0x7fffffff; }
    for (int q = lane; q < Q; q += 32) {
      float x[12];
      float mx = -INFINITY;
      int am = 0;
#pragma unroll
      for (int c = 0; c < 12; ++c) {
        x[c] = lg[q * 12 + c];
        if (x[c] > mx) { mx = x[c]; am = c; }  // first maximum, like np.argmax
      }
      float inv = 1.0f;
      if (!d.post_processed) {               // raw class logits: PostProcess' softmax
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 12; ++c) { x[c] = expf(x[c] - mx); sum += x[c]; }
        inv = 1.0f / sum;
      }                                       // else: PostProcess output, the probabilities ARE the scores
      const float score = x[am] * inv;
      const long long gq = static_cast<long long>(img) * Q + q;
      if (d.probs) {
#pragma unroll
        for (int c = 0; c < 12; ++c) d.probs[gq * 12 + c] = x[c] * inv;
      }
      if (d.points_px) {
        // fp32 multiply then add, unfused, exactly like `pt[:, 0] * width + x1` on float32 tensors
        d.points_px[gq * 2 + 0] = __fadd_rn(__fmul_rn(pt[q * 2 + 0], bw), bx1);
        d.points_px[gq * 2 + 1] = __fadd_rn(__fmul_rn(pt[q * 2 + 1], bh), by1);
      }
      if (d.sigmas && d.logsig) {
        d.sigmas[gq * 2 + 0] = expf(d.logsig[gq * 2 + 0]);
        d.sigmas[gq * 2 + 1] = expf(d.logsig[gq * 2 + 1]);
      }
      if (am != 11) {
#pragma unroll
        for (int l = 0; l < 11; ++l)
          if (l == am && score > best_s[l]) { best_s[l] = score; best_q[l] = q; }  // q ascending: first max kept
      }
    }
    // warp arg-max per label: higher score wins, ties go to the lower query index
    int n = 0;
#pragma unroll
    for (int l = 0; l < 11; ++l) {
      float sc = best_s[l];
      int qi = best_q[l];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float s2 = __shfl_xor_sync(FULL, sc, o);
        const int q2 = __shfl_xor_sync(FULL, qi, o);
        if (s2 > sc || (s2 == sc && q2 < qi)) { sc = s2; qi = q2; }
      }
      const bool present = sc >= 0.f;
      if (lane == 0) {
        d.assign[img * 11 + l] = present ? qi : -1;
        if (present) {
          s_lab[n] = l;
          const float px = __fadd_rn(__fmul_rn(pt[qi * 2 + 0], bw), bx1);
          const float py = __fadd_rn(__fmul_rn(pt[qi * 2 + 1], bh), by1);
          s_uv[2 * n] = static_cast<double>(px);
          s_uv[2 * n + 1] = static_cast<double>(py);
          // bearing of the correspondence
          const double bx = (static_cast<double>(px) - kCx) / kFx, by = (static_cast<double>(py) - kCy) / kFy;
          const double inv = 1.0 / sqrt(bx * bx + by * by + 1.0);
          s_bear[3 * n] = bx * inv; s_bear[3 * n + 1] = by * inv; s_bear[3 * n + 2] = inv;
          if (d.logsig) {
            s_sig[2 * n] = static_cast<double>(expf(d.logsig[(static_cast<long long>(img) * Q + qi) * 2 + 0]));
            s_sig[2 * n + 1] = static_cast<double>(expf(d.logsig[(static_cast<long long>(img) * Q + qi) * 2 + 1]));
          } else {
            s_sig[2 * n] = 1.0; s_sig[2 * n + 1] = 1.0;
          }
        }
      }
      if (present) ++n;
    }
    if (lane == 0) s_n = n;
  }
  __syncthreads();
  const int n = s_n;

  double quat[4] = {0, 0, 0, 0}, tv[3] = {0, 0, 0};
  unsigned inl_mask = 0u;
  auto write_out = [&](int st) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < 4; ++i) d.quat[img * 4 + i] = quat[i];
      for (int i = 0; i < 3; ++i) d.tvec[img * 3 + i] = tv[i];
      d.status[img] = st;
      if (d.inlier_mask) d.inlier_mask[img] = static_cast<int32_t>(inl_mask);
    }
  };
  if (n < 4) { write_out(1); return; }  // cv2.solvePnPRansac raises -> caller records the zero pose

  const long long t_assign = clock64();
  // ---- exhaustive minimal-sample consensus: triple #c goes to thread c % kPnpThreads
  Hyp best;
  best.cnt = 0; best.err = 1e300; best.mask = 0u;
#pragma unroll
  for (int i = 0; i < 9; ++i) best.R[i] = 0.0;
  best.t[0] = best.t[1] = best.t[2] = 0.0;
  const double thr = static_cast<double>(d.reproj_dev ? d.reproj_dev[img] : d.reproj_thresh);
  // exactly four correspondences: cv2.solvePnPRansac skips RANSAC (model_points == npoints) and calls solvePnP(P3P),
  // which solves the FIRST three correspondences and lets the fourth pick among the (up to four) solutions; all four
  // are reported as inliers and no threshold is applied
  const double thr2 = n == 4 ? INFINITY : thr * thr;
  // Every thread first decodes its own triple index, then all lanes run the solver together: calling it from inside
  // the enumeration loop would serialise the warp (one active lane per iteration).
  const int ntriples = n == 4 ? 1 : n * (n - 1) * (n - 2) / 6;     // n == 4: triple #0 = correspondences (0, 1, 2)
  for (int c = static_cast<int>(threadIdx.x); c < ((ntriples + kPnpThreads - 1) / kPnpThreads) * kPnpThreads;
       c += kPnpThreads) {
    int i0 = 0, i1 = 1, i2 = 2;
    if (c < ntriples) {
      int rem = c;
      for (i0 = 0; i0 < n - 2; ++i0) {                    // triples starting with i0: C(n-1-i0, 2)
        const int cnt0 = (n - 1 - i0) * (n - 2 - i0) / 2;
        if (rem < cnt0) break;
        rem -= cnt0;
      }
      for (i1 = i0 + 1; i1 < n - 1; ++i1) {               // pairs (i1, i2) with i2 > i1: n-1-i1 each
        const int cnt1 = n - 1 - i1;
        if (rem < cnt1) break;
        rem -= cnt1;
      }
      i2 = i1 + 1 + rem;
    }
    Hyp cand;
    cand.cnt = 0; cand.err = 1e300; cand.mask = 0u;
    p3p_consensus(i0, i1, i2, n, s_lab, s_uv, s_bear, thr2, cand);   // lanes beyond ntriples redo triple (0,1,2)
    if (c < ntriples && (cand.cnt > best.cnt || (cand.cnt == best.cnt && cand.cnt > 0 && cand.err < best.err)))
      best = cand;
  }
  __syncwarp();
  // arg-best: more inliers, then lower error, then lower thread index
  int bl = lane, bc = best.cnt;
  double be = best.err;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int c2 = __shfl_xor_sync(FULL, bc, o);
    const double e2 = __shfl_xor_sync(FULL, be, o);
    const int l2 = __shfl_xor_sync(FULL, bl, o);
    if (c2 > bc || (c2 == bc && (e2 < be || (e2 == be && l2 < bl)))) { bc = c2; be = e2; bl = l2; }
  }
  if (lane == bl) {
    s_bcnt[warp] = best.cnt; s_berr[warp] = best.err; s_bmask[warp] = best.mask;
#pragma unroll
    for (int i = 0; i < 9; ++i) s_bpose[warp][i] = best.R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) s_bpose[warp][9 + i] = best.t[i];
  }
  __syncthreads();
  if (warp != 0) return;
  const long long t_cons = clock64();
  int bw_ = 0;
  for (int w = 1; w < kPnpWarps; ++w)
    if (s_bcnt[w] > s_bcnt[bw_] || (s_bcnt[w] == s_bcnt[bw_] && s_berr[w] < s_berr[bw_])) bw_ = w;
  if (s_bcnt[bw_] < 4) { write_out(1); return; }  // no hypothesis supported by >= 4 correspondences
  double R[9], t[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = s_bpose[bw_][i];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = s_bpose[bw_][9 + i];
  inl_mask = s_bmask[bw_];

  // ---- refinement on the inliers
  PointObs o;
  o.active = (lane < n) && ((inl_mask >> lane) & 1u);
  o.X[0] = o.X[1] = o.X[2] = 0.0; o.mu = o.mv = 0.0; o.wu = o.wv = 1.0;
  if (lane < n) {
    const double* X = c_world + s_lab[lane] * 3;
    o.X[0] = X[0]; o.X[1] = X[1]; o.X[2] = X[2];
  }
  bool ok;
  if (d.weighted && d.logsig) {
    // SA/utils/speed_eval.py:285-288: w = 1/(sqrt(sigma)+1e-6), normalised to sum 1 per axis over the inliers
    double wu = 0.0, wv = 0.0;
    if (o.active) { wu = 1.0 / (sqrt(s_sig[2 * lane]) + 1e-6); wv = 1.0 / (sqrt(s_sig[2 * lane + 1]) + 1e-6); }
    const double su = wsum(wu), sv = wsum(wv);
    o.wu = wu / su; o.wv = wv / sv;
    if (lane < n) { o.mu = (s_uv[2 * lane] - kCx) / kFx; o.mv = (s_uv[2 * lane + 1] - kCy) / kFy; }
    ok = lm_refine(R, t, o, 1.0, 0.0, 1.0, 0.0, 0.005, n, s_J, s_A);
  } else {
    if (lane < n) { o.mu = s_uv[2 * lane]; o.mv = s_uv[2 * lane + 1]; }
    ok = lm_refine(R, t, o, kFx, kCx, kFy, kCy, 0.0, n, s_J, s_A);
  }
  if (!ok) { inl_mask = 0u; write_out(2); return; }
  if (d.debug_timing && img < 4 && lane == 0)
    printf("[spe pnp] img %d n=%d cycles: assign %lld consensus %lld lm %lld\n", img, n, t_assign - t_start,
           t_cons - t_assign, clock64() - t_cons);

  // ---- self-assessment statistics: inlier RMS reprojection error (pixels), mean predicted sigma (pixels)
  double e2 = 0.0, sg = 0.0;
  if (o.active) {
    const double x = R[0] * o.X[0] + R[1] * o.X[1] + R[2] * o.X[2] + t[0];
    const double y = R[3] * o.X[0] + R[4] * o.X[1] + R[5] * o.X[2] + t[1];
    const double z = R[6] * o.X[0] + R[7] * o.X[1] + R[8] * o.X[2] + t[2];
    const double du = kFx * x / z + kCx - s_uv[2 * lane], dv = kFy * y / z + kCy - s_uv[2 * lane + 1];
    e2 = du * du + dv * dv;
    sg = 0.5 * (s_sig[2 * lane] + s_sig[2 * lane + 1]);
  }
  const int ninl = __popc(inl_mask);
  const double rms = sqrt(wsum(e2) / ninl);
  // sigma is in normalised crop units: the crop side converts it to pixels (callers that hand over pixel keypoints
  // pass the side separately; without it the sigma criterion cannot be evaluated and is skipped)
  const double side = d.post_processed ? static_cast<double>(d.sigma_px_scale) : static_cast<double>(bw);
  const double mean_sigma_px = wsum(sg) / ninl * side;
  int st = 0;
  if (d.reject) {
    const bool rej = (ninl < 4) || (rms > static_cast<double>(d.reject_rms_px)) ||
                     (d.logsig != nullptr && side > 0.0 && mean_sigma_px > static_cast<double>(d.reject_sigma));
    if (rej) st = 3;
  }
  rot_to_quat(R, quat);
  tv[0] = t[0]; tv[1] = t[1]; tv[2] = t[2];
  write_out(st);
}

// speed_score, RV/utils/speed_eval.py:245-262
__global__ void speed_score_kernel(const double* __restrict__ q_pr, const double* __restrict__ t_pr,
                                   const double* __restrict__ q_gt, const double* __restrict__ t_gt, int B,
                                   double* __restrict__ s_t, double* __restrict__ s_q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const double sp = q_pr[4 * i] < 0 ? -1.0 : 1.0, sg = q_gt[4 * i] < 0 ? -1.0 : 1.0;
  double dot = 0.0, dn = 0.0, gn = 0.0;
  for (int k = 0; k < 4; ++k) dot += (sp * q_pr[4 * i + k]) * (sg * q_gt[4 * i + k]);
  for (int k = 0; k < 3; ++k) {
    const double e = t_pr[3 * i + k] - t_gt[3 * i + k];
    dn += e * e;
    gn += t_gt[3 * i + k] * t_gt[3 * i + k];
  }
  s_t[i] = sqrt(dn) / sqrt(gn);
  s_q[i] = 2.0 * acos(fmin(fabs(dot), 1.0));
}

}  // namespace

std::string launch_speed_score(const double* q_pr, const double* t_pr, const double* q_gt, const double* t_gt, int B,
                               double* s_t, double* s_q, cudaStream_t s) {
  if (B <= 0) return "";
  ProfScope ps(kFamPnp, s);
  speed_score_kernel<<<(B + 127) / 128, 128, 0, s>>>(q_pr, t_pr, q_gt, t_gt, B, s_t, s_q);
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

std::string launch_assign_pnp(const PnpDesc& d, cudaStream_t s) {
  if (d.B <= 0) return "";
  if (d.Q <= 0 || d.Q > 4096) return "assign_pnp: bad query count";
  if (d.num_models < 0 || static_cast<long long>(d.num_models) * d.Q > kMaxPooled)
    return "ensemble_pnp: models x queries must not exceed 4096";
  ProfScope ps(kFamPnp, s);
  PnpDesc dd = d;
  static const bool timing = getenv("SPE_PNP_TIMING") != nullptr;
  dd.debug_timing = timing ? 1 : 0;
  SPE_CUDA_TRY(launch_pdl(assign_pnp_kernel, dim3(d.B), dim3(kPnpThreads), 0, s, dd));
  SPE_CUDA_TRY(cudaGetLastError());
  return "";
}

}  // namespace spe
