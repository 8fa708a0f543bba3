// detmatch_b200/csrc/pib_dev.cuh -- device pieces shared by the point-in-box kernels
// (points_in_boxes.cu) and RoI-aware pooling (roiaware_pool3d.cu): glibc-exact sinf / cosf, the
// prepared box and the exact inside test of points_in_boxes_cpu.cpp:16-40.  Included INSIDE the
// including file's anonymous namespace (the __constant__ tables get internal linkage per file).
#pragma once

// ------------------------------------------------------------------------------------------
// glibc >= 2.28 sinf/cosf (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h), restated.
// oracle/pcfe_oracle.c carries the same restatement for the CPU and is checked against the
// host libm exhaustively on [2^-14, 120) (tests/test_oracle.py).
// ------------------------------------------------------------------------------------------
struct SinCosTab {
  double c0, c1, c2, c3, c4, s1, s2, s3;
};
__device__ __constant__ SinCosTab kTab[2] = {
    {0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10,
     0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
    {-0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10,
     -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
__device__ __constant__ uint32_t kInvPio4[24] = {
    0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44,
    0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1,
    0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62,
    0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};

__device__ __forceinline__ uint32_t abstop12(float f) { return (__float_as_uint(f) >> 20) & 0x7ffu; }

// n even: sine polynomial, n odd: cosine polynomial.  Multiplications are single-rounded
// (__dmul_rn), every a + b*c is ONE fused operation (__fma_rn), matching glibc's FMA build.
__device__ __forceinline__ float sc_poly(double x, double x2, const SinCosTab& p, int n) {
  if ((n & 1) == 0) {
    const double x3 = __dmul_rn(x, x2);
    const double s1 = __fma_rn(x2, p.s3, p.s2);
    const double x7 = __dmul_rn(x3, x2);
    const double s = __fma_rn(x3, p.s1, x);
    return __double2float_rn(__fma_rn(x7, s1, s));
  } else {
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __fma_rn(x2, p.c4, p.c3);
    const double c1 = __fma_rn(x2, p.c1, p.c0);
    const double x6 = __dmul_rn(x4, x2);
    const double c = __fma_rn(x4, p.c2, c1);
    return __double2float_rn(__fma_rn(x6, c2, c));
  }
}

__device__ float glibc_sin_or_cos(float y, int want_cos) {
  double x = (double)y;
  const double sign[4] = {1.0, -1.0, -1.0, 1.0};
  int n;
  int tab = 0;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    if (abstop12(y) < abstop12(0x1p-12f)) return want_cos ? 1.0f : y;
    return sc_poly(x, __dmul_rn(x, x), kTab[0], want_cos);
  } else if (abstop12(y) < abstop12(120.0f)) {
    const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
    n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __fma_rn(-(double)n, 0x1.921FB54442D18p0, x);
    const double s = sign[n & 3];
    if (n & 2) tab = 1;
    return sc_poly(__dmul_rn(x, s), __dmul_rn(x, x), kTab[tab], n ^ want_cos);
  } else if (abstop12(y) < 0x7f8u) {
    uint32_t xi = __float_as_uint(y);
    const int sgn = (int)(xi >> 31);
    const uint32_t* arr = &kInvPio4[(xi >> 26) & 15];
    const int shift = (xi >> 23) & 7;
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    uint64_t res0 = (uint64_t)(uint32_t)(xi * arr[0]);
    const uint64_t res1 = (uint64_t)xi * arr[4];
    const uint64_t res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    const uint64_t nn = (res0 + (1ULL << 61)) >> 62;
    res0 -= nn << 62;
    x = __dmul_rn((double)(int64_t)res0, 0x1.921FB54442D18p-62);
    n = (int)nn;
    const double s = sign[(n + sgn) & 3];
    if ((n + sgn) & 2) tab = 1;
    return sc_poly(__dmul_rn(x, s), __dmul_rn(x, x), kTab[tab], n ^ want_cos);
  }
  return __fsub_rn(y, y);  // Inf/NaN -> NaN
}

// ------------------------------------------------------------------------------------------
// prepared boxes
// ------------------------------------------------------------------------------------------
struct __align__(16) PBox {
  float cx, cy, czc, hh;      // centre (z shifted to the box centre), half height
  float cosa, sina, hl, hw;   // rotation by rz + pi/2, half length (local x), half width (local y)
};
static_assert(sizeof(PBox) == 32, "PBox must be 32 bytes");

// Conservative reject data: a point with |x - cx| > r or |y - cy| > r is outside whatever the
// rotation.  r = 1.0001 * sqrt(hl^2 + hw^2) rounded up: an inside point has computed
// lx^2 + ly^2 < hl^2 + hw^2, and the float rotation changes the norm of (sx, sy) by less than 1e-6
// relative, so |sx| > r or |sy| > r implies the exact test fails.  NaN compares false (no reject),
// r = NaN / Inf never rejects: the exact test decides.  99.5 % of the point-box pairs of a LiDAR
// frame are rejected with 2 subtractions and 2 comparisons instead of the 20-instruction test.
struct __align__(16) RBox {
  float cx, cy, r, pad;
};
static_assert(sizeof(RBox) == 16, "RBox must be 16 bytes");

// mmdet3d box (cx, cy, cz_bottom, w, l, h, rz) -> prepared box (points_in_boxes_cpu.cpp:16-40)
__device__ __forceinline__ PBox make_pbox_mmdet(const float* __restrict__ b) {
  const float cx = b[0], cy = b[1], cz = b[2], w = b[3], l = b[4], h = b[5], rz = b[6];
  PBox p;
  p.cx = cx;
  p.cy = cy;
  // cz += h / 2.0  (double add, rounded to float; points_in_boxes_cpu.cpp:33)
  p.czc = __double2float_rn(__dadd_rn((double)cz, __dmul_rn((double)h, 0.5)));
  // The reference compares float values against the DOUBLE h/2, l/2, w/2 (:35,:37-38).  x/2 is
  // exact in float except for odd subnormals; directed rounding keeps the float comparison
  // equivalent there too:  |dz| > h/2  <=>  |dz| > rd(h/2);   lx < l/2  <=>  lx < ru(l/2);
  // lx > -l/2  <=>  lx > -ru(l/2).
  p.hh = __double2float_rd(__dmul_rn((double)h, 0.5));
  p.hl = __double2float_ru(__dmul_rn((double)l, 0.5));
  p.hw = __double2float_ru(__dmul_rn((double)w, 0.5));
  // rot_angle = rz + M_PI / 2  (double add, rounded to float; :19)
  const float rot = __double2float_rn(__dadd_rn((double)rz, 0x1.921fb54442d18p+0));
  p.cosa = glibc_sin_or_cos(rot, 1);
  p.sina = glibc_sin_or_cos(rot, 0);
  return p;
}

// reject record of a prepared box (see RBox)
__device__ __forceinline__ RBox make_rbox(const PBox& p) {
  RBox r;
  r.cx = p.cx;
  r.cy = p.cy;
  const double hl = (double)p.hl, hw = (double)p.hw;
  r.r = __double2float_ru(__dmul_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(hl, hl), __dmul_rn(hw, hw))), 1.0001));
  r.pad = 0.0f;
  return r;
}

// true: the pair cannot be inside (see RBox); false: run the exact test
__device__ __forceinline__ bool xy_reject(float x, float y, const RBox& r) {
  return (fabsf(__fsub_rn(x, r.cx)) > r.r) | (fabsf(__fsub_rn(y, r.cy)) > r.r);
}

// points_in_boxes_cpu.cpp:25-40, float32, no contraction.
__device__ __forceinline__ int in_box(float x, float y, float z, const PBox& b) {
  const float dz = __fsub_rn(z, b.czc);
  const bool z_out = fabsf(dz) > b.hh;  // reject-if-greater keeps the NaN-z asymmetry
  const float sx = __fsub_rn(x, b.cx);
  const float sy = __fsub_rn(y, b.cy);
  const float lx = __fadd_rn(__fmul_rn(sx, b.cosa), __fmul_rn(sy, -b.sina));
  const float ly = __fadd_rn(__fmul_rn(sx, b.sina), __fmul_rn(sy, b.cosa));
  // (lx > -hl) & (lx < hl)  <=>  |lx| < hl for every hl (negative, zero, Inf and NaN included:
  // both forms are false whenever hl <= 0 or anything is NaN, and |+-Inf| < Inf is false like
  // the two one-sided tests)
  const bool in = (fabsf(lx) < b.hl) & (fabsf(ly) < b.hw);
  return (in & !z_out) ? 1 : 0;
}

