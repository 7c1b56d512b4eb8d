/*
 * oracle/cpu_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP, strict fp32, rounding defined below) of the
 * differentiable ray-march in the reference's differender/volume_raycaster.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build, load or call this file.  The product path
 * (differender_b200/) never does.
 *
 * PARITY: PINNED AGAINST THE REFERENCE'S SOURCE, UNPINNED AGAINST ITS COMPILER.  The reference has no tests, golden
 * vectors or fixtures, and its arithmetic is JIT-compiled by the un-vendored, unpinned third-party `taichi` and
 * `taichi_glsl` packages (requirements.txt:1, setup.py:22-23), which are not installed here and cannot be (no network).
 *   (a) Algorithm: oracle/ti_shim.py is a strict-IEEE-fp32 interpreter of the Taichi subset the reference uses; with it
 *       the reference's OWN SOURCE (read from /root/reference, never copied) runs in this container.  The build of
 *       this file with every contraction switched off (`source_order`, oracle/cpu_oracle.py VARIANTS) agrees with it
 *       BIT FOR BIT on image, sample counts and early-termination counts, and to <= 2e-6 relative L2 on both gradients
 *       (forward + backward at three sampling rates, jittered and not, non-cubic volume, camera inside the box, another
 *       frustum, flat block, nondiff path; and through the reference's public Raycaster API, batched and not):
 *       tests/test_shim_pin.py, tests/golden/shim/ (+ make_shim_golden.py), profiles/r02_shim_pin_report.txt.  The two
 *       deliberate deviations are SURVEY 7.3 H3 (n == 1 rays: 0/0 sample position in the reference, t = 0 here) and H4
 *       (the reference NaN-poisons and then zeroes the gradient of voxels under an exactly flat sample; kept finite here).
 *   (b) Rounding: what the real Taichi compiler contracts or approximates (fast_math) is NOT pinned; this file defines
 *       those choices (below), tools/rounding_envelope.py measures how far the alternatives move the results, and
 *       oracle/taichi_probe.py compares against real Taichi on the first box that has it.
 *   (c) The primitives (mix, clamp, cross, reflect, normalized, ti.floor/pow/min/max, the adjoints of max/min) are
 *       restated from the packages' published definitions, here and in the interpreter alike; further cross-checks:
 *       oracle/torch_ref.py (an independent fp64 PyTorch restatement differentiated by torch.autograd), fp64 finite
 *       differences and analytic known-answer tests (tests/test_oracle_*.py).
 *
 * ROUNDING IS DEFINED HERE.  The normal is a central difference over +-1e-3 world units
 * (:191-203), i.e. a catastrophic cancellation whose result depends on how each trilinear tap is
 * rounded.  Taichi's default fast_math=True lets LLVM contract `a*(1-t) + b*t`;
 * this oracle fixes that choice explicitly so that two implementations can agree bit for bit on
 * the ill-conditioned part:
 *     mix(a,b,t)      := fma(a, 1-t, b*t)         pos := fma(t, dir, cam)
 *     0.5*p + 0.5     := fma(0.5, p, 0.5)         A_s := fma(1 - A_{s-1}.w, C, A_{s-1})
 *     float(s)/float(n-1) := s * fl(1/(n-1))      (reciprocal-multiply, what fast-math emits for a division)
 * Everything else is one IEEE-rounded operation per source operator, in source order
 * (-ffp-contract=off).  With -DORACLE_FP64 the same code runs in double (used to pin the adjoint
 * against torch.autograd in float64).
 *
 * ROUNDING ENVELOPE.  Each choice above (and a few more that Taichi's fast_math leaves open) has a compile-time
 * switch that selects another plausible rounding; tools/rounding_envelope.py builds every variant and measures how far
 * it moves the image and the gradients from the default build (profiles/r02_rounding_envelope.txt):
 *     -DORA_MIX_UNFUSED        mix := fl(fl(a*(1-t)) + fl(b*t))           (no contraction)
 *     -DORA_MIX_FMA_B          mix := fma(b, t, fl(a*(1-t)))              (the other contraction)
 *     -DORA_DIV_TRUE           float(s)/float(n-1) as one IEEE division   (:279-280)
 *     -DORA_DIV_APPROX         s * r, r = reciprocal off by one ulp on odd n (div.approx-style, <= 2 ulp)
 *     -DORA_POW_FAST           pow(x, y) := exp2f(y * log2f(x))           (libdevice fast pow, :284-285, :296)
 *     -DORA_TAN_F32            near_h / near_w folded in fp32 with tanf   (:146-147)
 *     -DORA_POS_UNFUSED        pos := cam + fl(t*dir), 0.5*p + 0.5 unfused (:163-165, :277)
 *     -DORA_COMPOSITE_UNFUSED  A_s := fl(fl((1-A.w)*C) + A)               (:300-302)
 *     -DORA_NORMALIZE_DIV      normalized() := v / |v| per component      (:203, :290)
 * None of them is used by a parity test: the default build IS the oracle.
 *
 * Every function cites the reference lines it follows
 * (paths relative to /root/reference/differender/volume_raycaster.py).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_FP64
typedef double real;
#define R_SQRT sqrt
#define R_FLOOR floor
#define R_POW pow
#define R_FMAX fmax
#define R_FMIN fmin
#define R_FMA fma
#else
typedef float real;
#define R_SQRT sqrtf
#define R_FLOOR floorf
#define R_POW powf
#define R_FMAX fmaxf
#define R_FMIN fminf
#define R_FMA fmaf
#endif
#define RC(x) ((real)(x))

typedef struct {
    /* volume dims in the reference's Taichi order (volume_raycaster.py:481):
     * X = torch W (unit stride), Y = torch D (slowest), Z = torch H. */
    int X, Y, Z;
    int W, H;          /* render resolution (w, h)                    :74      */
    int R;             /* tf resolution                               :113     */
    int M;             /* max_samples                                 :90      */
    double sr;         /* sampling rate                               :222,262 */
    double fov_deg;    /* :76-77 */
    double near_;      /* :78    */
    int nondiff;       /* 1: raycast_nondiff + get_final_image_nondiff :308-361 */
    int has_jitter;    /* jitter flag of compute_entry_exit           :254     */
} OraDesc;

/* constants folded on the host in double, then rounded to fp32 (Python scalars
 * captured by the Taichi kernels: :75-78, 146-147, 165, 215, 248-249) */
typedef struct {
    real near_, near_w, near_h;
    real scale[3];
    real vol_diag;
    real tf_len;
    real sr, inv_sr;
    int dim[3];
} OraConst;

static void fold_constants(const OraDesc *d, OraConst *c)
{
    const double pi = 3.14159265358979323846;
    double fov_rad = d->fov_deg * (pi / 180.0);                 /* np.radians :77 */
    double aspect = (double)d->W / (double)d->H;                /* :75  */
    double near_h = 2.0 * tan(fov_rad) * d->near_;              /* :146 */
    double near_w = near_h * aspect;                            /* :147 */
    c->near_ = (real)(float)d->near_;
    c->near_h = (real)(float)near_h;
    c->near_w = (real)(float)near_w;
#if defined(ORA_TAN_F32) && !defined(ORACLE_FP64)
    c->near_h = 2.0f * tanf((float)fov_rad) * (float)d->near_;  /* evaluated in the kernel's fp32 instead of folded in double */
    c->near_w = c->near_h * (float)aspect;
#endif
    c->dim[0] = d->X; c->dim[1] = d->Y; c->dim[2] = d->Z;
    for (int a = 0; a < 3; ++a)
        c->scale[a] = (real)(float)((double)c->dim[a] - 1.0 - 1e-4);  /* :165 */
    double dx = d->X - 1.0, dy = d->Y - 1.0, dz = d->Z - 1.0;
    c->vol_diag = (real)(float)sqrt(dx * dx + dy * dy + dz * dz);     /* :248-249 */
    c->tf_len = (real)(float)(d->R - 1);                              /* :215 */
    c->sr = (real)(float)d->sr;
    c->inv_sr = (real)(float)(1.0 / d->sr);                           /* :285 (1.0 / sampling_rate) */
}

typedef struct { real x, y, z; } v3;
typedef struct { real x, y, z, w; } v4;

/* taichi_glsl.mix(x, y, a) = x*(1-a) + y*a, contracted as LLVM does: fma(x, 1-a, y*a) (see header) */
static inline real mixf(real a, real b, real t)
{
#if defined(ORA_MIX_UNFUSED)
    real p = a * (RC(1.0) - t), q = b * t;
    return p + q;
#elif defined(ORA_MIX_FMA_B)
    return R_FMA(b, t, a * (RC(1.0) - t));
#else
    return R_FMA(a, RC(1.0) - t, b * t);
#endif
}
/* a*b + c as written at :163-165 (0.5*pos + 0.5), :277 (look_from + t*vd) */
static inline real muladd_pos(real a, real b, real c)
{
#if defined(ORA_POS_UNFUSED)
    real p = a * b;
    return p + c;
#else
    return R_FMA(a, b, c);
#endif
}
static inline real pow_r(real x, real y)
{
#if defined(ORA_POW_FAST) && !defined(ORACLE_FP64)
    return exp2f(y * log2f(x));
#else
    return R_POW(x, y);
#endif
}
static inline real dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross3(v3 a, v3 b)
{
    v3 r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
    return r;
}
/* Taichi Vector.normalized(eps=0): invlen = 1/(norm + eps); v * invlen */
static inline v3 normalized3(v3 a)
{
#if defined(ORA_NORMALIZE_DIV)
    real len = R_SQRT(dot3(a, a));
    v3 r = { a.x / len, a.y / len, a.z / len };
#else
    real inv = RC(1.0) / R_SQRT(dot3(a, a));
    v3 r = { a.x * inv, a.y * inv, a.z * inv };
#endif
    return r;
}

/* low_high_frac  :7-21 */
static inline void low_high_frac(real x, int *lo, int *hi, real *frac)
{
    x = R_FMAX(x, RC(0.0));
    real l = R_FLOOR(x);
    *lo = (int)l;
    *hi = (int)(l + RC(1.0));
    *frac = x - l;
}

/* get_entry_exit_points  :28-53  (box = [-1,1]^3, :235-238) */
static inline int entry_exit(v3 o, v3 d, real *tmin_o, real *tmax_o)
{
    real ix = RC(1.0) / d.x, iy = RC(1.0) / d.y, iz = RC(1.0) / d.z;
    real t1 = (-RC(1.0) - o.x) * ix, t2 = (RC(1.0) - o.x) * ix;
    real t3 = (-RC(1.0) - o.y) * iy, t4 = (RC(1.0) - o.y) * iy;
    real t5 = (-RC(1.0) - o.z) * iz, t6 = (RC(1.0) - o.z) * iz;
    real tmin = R_FMAX(R_FMAX(R_FMIN(t1, t2), R_FMIN(t3, t4)), R_FMIN(t5, t6));
    real tmax = R_FMIN(R_FMIN(R_FMAX(t1, t2), R_FMAX(t3, t4)), R_FMAX(t5, t6));
    *tmin_o = tmin; *tmax_o = tmax;
    return !(tmax < RC(0.0) || tmin > tmax);
}

/* get_ray_direction  :127-151 */
static inline v3 ray_direction(const OraConst *c, v3 orig, v3 view, real x, real y)
{
    real u = x - RC(0.5), v = y - RC(0.5);
    v3 up0 = { RC(0.0), RC(1.0), RC(0.0) };
    v3 right = normalized3(cross3(view, up0));
    v3 up = normalized3(cross3(right, view));
    v3 near_m = { orig.x + c->near_ * view.x, orig.y + c->near_ * view.y, orig.z + c->near_ * view.z };
    real a = u * c->near_w, b = v * c->near_h;
    v3 np_ = { near_m.x + a * right.x + b * up.x,
               near_m.y + a * right.y + b * up.y,
               near_m.z + a * right.z + b * up.z };
    v3 dd = { np_.x - orig.x, np_.y - orig.y, np_.z - orig.z };
    return normalized3(dd);
}

/* torch layout of the volume: vol[y][z][x]  (:481, 566, 571) */
static inline size_t vidx(const OraConst *c, int x, int y, int z)
{
    return ((size_t)y * (size_t)c->dim[2] + (size_t)z) * (size_t)c->dim[0] + (size_t)x;
}

typedef struct { int x0, x1, y0, y1, z0, z1; real fx, fy, fz; } Cell;

/* address part of sample_volume_trilinear  :163-172 */
static inline void locate(const OraConst *c, v3 pos, Cell *k)
{
    real px = R_FMIN(RC(1.0), R_FMAX(RC(0.0), muladd_pos(RC(0.5), pos.x, RC(0.5)))) * c->scale[0];
    real py = R_FMIN(RC(1.0), R_FMAX(RC(0.0), muladd_pos(RC(0.5), pos.y, RC(0.5)))) * c->scale[1];
    real pz = R_FMIN(RC(1.0), R_FMAX(RC(0.0), muladd_pos(RC(0.5), pos.z, RC(0.5)))) * c->scale[2];
    low_high_frac(px, &k->x0, &k->x1, &k->fx);
    low_high_frac(py, &k->y0, &k->y1, &k->fy);
    low_high_frac(pz, &k->z0, &k->z1, &k->fz);
    if (k->x1 > c->dim[0] - 1) k->x1 = c->dim[0] - 1;
    if (k->y1 > c->dim[1] - 1) k->y1 = c->dim[1] - 1;
    if (k->z1 > c->dim[2] - 1) k->z1 = c->dim[2] - 1;
}

/* interpolation part of sample_volume_trilinear  :173-189 */
static inline real trilinear(const OraConst *c, const real *vol, v3 pos, Cell *k)
{
    locate(c, pos, k);
    real v000 = vol[vidx(c, k->x0, k->y0, k->z0)], v100 = vol[vidx(c, k->x1, k->y0, k->z0)];
    real v010 = vol[vidx(c, k->x0, k->y1, k->z0)], v110 = vol[vidx(c, k->x1, k->y1, k->z0)];
    real v001 = vol[vidx(c, k->x0, k->y0, k->z1)], v101 = vol[vidx(c, k->x1, k->y0, k->z1)];
    real v011 = vol[vidx(c, k->x0, k->y1, k->z1)], v111 = vol[vidx(c, k->x1, k->y1, k->z1)];
    real a = mixf(v000, v100, k->fx), b = mixf(v010, v110, k->fx);
    real lo = mixf(a, b, k->fy);
    a = mixf(v001, v101, k->fx); b = mixf(v011, v111, k->fx);
    real hi = mixf(a, b, k->fy);
    return mixf(lo, hi, k->fz);
}

/* reverse of :173-189: scatter `adj` with the 8 trilinear weights */
static inline void trilinear_adjoint(const OraConst *c, double *gvol, const Cell *k, real adj)
{
    real lo = adj * (RC(1.0) - k->fz), hi = adj * k->fz;
    real lo_a = lo * (RC(1.0) - k->fy), lo_b = lo * k->fy;
    real hi_a = hi * (RC(1.0) - k->fy), hi_b = hi * k->fy;
    real gx0 = RC(1.0) - k->fx, gx1 = k->fx;
    const size_t id[8] = {
        vidx(c, k->x0, k->y0, k->z0), vidx(c, k->x1, k->y0, k->z0),
        vidx(c, k->x0, k->y1, k->z0), vidx(c, k->x1, k->y1, k->z0),
        vidx(c, k->x0, k->y0, k->z1), vidx(c, k->x1, k->y0, k->z1),
        vidx(c, k->x0, k->y1, k->z1), vidx(c, k->x1, k->y1, k->z1) };
    const real w[8] = { lo_a * gx0, lo_a * gx1, lo_b * gx0, lo_b * gx1,
                         hi_a * gx0, hi_a * gx1, hi_b * gx0, hi_b * gx1 };
    for (int i = 0; i < 8; ++i) {
        double add = (double)w[i];
#pragma omp atomic
        gvol[id[i]] += add;
    }
}

/* apply_transfer_function  :205-219 ; tf is [R][4] */
static inline v4 apply_tf(const OraConst *c, const real *tf, int R, real intensity,
                          int *lo_o, int *hi_o, real *f_o, real *x_o)
{
    real x = intensity * c->tf_len;
    int lo, hi; real f;
    low_high_frac(x, &lo, &hi, &f);
    if (hi > R - 1) hi = R - 1;
    if (lo > R - 1) lo = R - 1;          /* intensity > 1 would index past tf_tex; clamp (no-op on [0,1]) */
    const real *a = tf + 4 * (size_t)lo, *b = tf + 4 * (size_t)hi;
    v4 r = { mixf(a[0], b[0], f), mixf(a[1], b[1], f), mixf(a[2], b[2], f), mixf(a[3], b[3], f) };
    *lo_o = lo; *hi_o = hi; *f_o = f; *x_o = x;
    return r;
}

typedef struct {
    v3 cam, dir;
    real entry, exit_;
    int n;
} Ray;

/* compute_entry_exit  :221-259 ; jit = J(i,j) or 0 */
static inline void setup_ray(const OraDesc *d, const OraConst *c, v3 cam, int i, int j, real jit, Ray *r)
{
    real nrm = RC(1.0) / R_SQRT(dot3(cam, cam));
    v3 view = { -cam.x * nrm, -cam.y * nrm, -cam.z * nrm };          /* (-look_from).normalized() :233 */
    real x = ((real)i + RC(0.5)) / (real)d->W;                        /* :239 */
    real y = ((real)j + RC(0.5)) / (real)d->H;                        /* :240 */
    v3 vd = ray_direction(c, cam, view, x, y);
    real tmin, tmax;
    int hit = entry_exit(cam, vd, &tmin, &tmax);
    real ray_len = tmax - tmin;                                      /* :250 */
    real nf = (real)hit * (R_FLOOR(c->sr * ray_len * c->vol_diag) + RC(1.0));   /* :251-253 */
    int n = hit ? (int)nf : 0;
    if (d->has_jitter && n > 0)
        tmin += jit * ray_len / nf;                                   /* :254-255 */
    r->cam = cam; r->dir = vd; r->entry = tmin; r->exit_ = tmax; r->n = n;
}

typedef struct {
    v3 pos;
    Cell cc;               /* centre tap cell */
    real I;
    int lo, hi; real f, x;   /* tf lookup */
    v4 c;                  /* sample colour */
    real o;               /* opacity */
    v3 g; real glen2;     /* un-normalised central-difference gradient */
    v3 N, l;
    real nl, ndl, rv, rdv, pw, kraw, k;
    v4 C;                  /* shaded premultiplied colour */
} Sample;

/* sample position  :270-280  (n==1 => t = tmin') */
static inline v3 sample_pos(const Ray *r, int s)
{
    real ray_len = r->exit_ - r->entry;                              /* :272 */
    real t0 = r->entry + RC(0.5) * ray_len / (real)r->n;               /* :273-275 */
    real t;
    if (r->n > 1) {
#if defined(ORA_DIV_TRUE)
        real q = (real)s / (real)(r->n - 1);
#elif defined(ORA_DIV_APPROX) && !defined(ORACLE_FP64)
        float rcp = 1.0f / (float)(r->n - 1);
        if (r->n & 1) rcp = nextafterf(rcp, 2.0f);                     /* an approximate reciprocal: off by one ulp on odd n */
        real q = (real)s * rcp;
#else
        real q = (real)s * (RC(1.0) / (real)(r->n - 1));
#endif
        t = mixf(t0, r->exit_, q);                                    /* :277-280, see header */
    } else t = t0;
    v3 p = { muladd_pos(t, r->dir.x, r->cam.x), muladd_pos(t, r->dir.y, r->cam.y), muladd_pos(t, r->dir.z, r->cam.z) };
    return p;
}

/* body of raycast / raycast_nondiff  :281-299 / :329-347 */
static inline void shade_sample(const OraDesc *d, const OraConst *c, const real *vol, const real *tf,
                                const Ray *r, int s, Sample *q, int want_normal)
{
    q->pos = sample_pos(r, s);
    q->I = trilinear(c, vol, q->pos, &q->cc);                          /* :282 */
    q->c = apply_tf(c, tf, d->R, q->I, &q->lo, &q->hi, &q->f, &q->x); /* :283 */
    q->o = RC(1.0) - pow_r(RC(1.0) - q->c.w, c->inv_sr);                      /* :284-285 */
    if (!want_normal) return;
    /* get_volume_normal  :191-203 */
    const real delta = RC(1e-3);
    Cell t;
    v3 p = q->pos, a, b;
    a = p; a.x = p.x + delta; b = p; b.x = p.x - delta;
    q->g.x = trilinear(c, vol, a, &t) - trilinear(c, vol, b, &t);
    a = p; a.y = p.y + delta; b = p; b.y = p.y - delta;
    q->g.y = trilinear(c, vol, a, &t) - trilinear(c, vol, b, &t);
    a = p; a.z = p.z + delta; b = p; b.z = p.z - delta;
    q->g.z = trilinear(c, vol, a, &t) - trilinear(c, vol, b, &t);
    q->glen2 = dot3(q->g, q->g);
    q->N = normalized3(q->g);                                          /* 0/0 -> NaN on flat data (H4) */
    v3 lp = { r->cam.x + RC(0.0), r->cam.y + RC(1.0), r->cam.z + RC(0.0) };     /* :281 */
    v3 ld = { q->pos.x - lp.x, q->pos.y - lp.y, q->pos.z - lp.z };
    q->l = normalized3(ld);                                            /* :288-290 */
    q->nl = dot3(q->N, q->l);
    q->ndl = R_FMAX(q->nl, RC(0.0));                                       /* :291 (NaN -> 0) */
    real two_nl = RC(2.0) * q->nl;                                       /* reflect: I - 2*dot(N,I)*N */
    v3 rr = { q->l.x - two_nl * q->N.x, q->l.y - two_nl * q->N.y, q->l.z - two_nl * q->N.z };
    v3 mv = { -r->dir.x, -r->dir.y, -r->dir.z };
    q->rv = dot3(rr, mv);
    q->rdv = R_FMAX(q->rv, RC(0.0));                                       /* :295 */
    q->pw = pow_r(q->rdv, RC(32.0));                                       /* :296 */
    q->kraw = RC(0.8) * q->ndl + RC(0.3) * q->pw + RC(0.4);                     /* diffuse + specular + ambient */
    q->k = d->nondiff ? q->kraw : R_FMIN(RC(1.0), q->kraw);                /* :298 vs :345 */
    q->C.x = q->k * q->c.x * q->o * RC(1.0);                              /* :297-299 */
    q->C.y = q->k * q->c.y * q->o * RC(1.0);
    q->C.z = q->k * q->c.z * q->o * RC(1.0);
    q->C.w = q->o;
}

static inline v4 composite(v4 A, v4 C)
{
    real T = RC(1.0) - A.w;                                              /* :300-302 */
#if defined(ORA_COMPOSITE_UNFUSED)
    real px = T * C.x, py = T * C.y, pz = T * C.z, pw = T * C.w;
    v4 r = { px + A.x, py + A.y, pz + A.z, pw + A.w };
#else
    v4 r = { R_FMA(T, C.x, A.x), R_FMA(T, C.y, A.y), R_FMA(T, C.z, A.z), R_FMA(T, C.w, A.w) };
#endif
    return r;
}

/* march one ray; returns final A, number of active samples K; if tape != NULL stores A_{s-1} for s < K */
static inline v4 march(const OraDesc *d, const OraConst *c, const real *vol, const real *tf,
                       const Ray *r, int *K_o, v4 *tape)
{
    v4 A = { 0, 0, 0, 0 };                                             /* H1: tape[-1] = 0 */
    int K = 0;
    Sample q;
    for (int s = 0; s < r->n; ++s) {
        if (d->nondiff) {
            if (!(A.w < RC(0.99))) break;                                 /* :318 (later iterations are no-ops) */
            shade_sample(d, c, vol, tf, r, s, &q, 0);
            if (q.c.w > RC(1e-3)) {                                       /* :334 */
                shade_sample(d, c, vol, tf, r, s, &q, 1);
                A = composite(A, q.C);
                ++K;
            }
        } else {
            if (!(A.w < RC(0.99) && s < d->M)) break;                     /* :267-269; else-branch copies A forward :304-306 */
            shade_sample(d, c, vol, tf, r, s, &q, 1);
            if (tape) tape[s] = A;
            A = composite(A, q.C);
            ++K;                                                       /* :303 */
        }
    }
    *K_o = K;
    return A;
}

/* ---------------------------------------------------------------------------------------------
 * Forward.  RaycastFunction.forward for one item  :431-438  /  raycast_nondiff :515-523
 *   vol      [Y][Z][X] fp32 (torch (D,H,W) contiguous)
 *   tf       [R][4]
 *   cam      [3]
 *   jitter   [W][H] raw order J(i,j) or NULL
 *   out_rgba [W][H][4] raw order (output_rgba.to_torch())
 *   out_K    [W][H] active samples per ray (valid_sample_step_count - 1, :303,367), may be NULL
 *   out_n    [W][H] sample_step_nums (:259), may be NULL
 * ------------------------------------------------------------------------------------------- */
void ora_forward(const OraDesc *d, const real *vol, const real *tf, const real *cam3,
                 const real *jitter, real *out_rgba, int32_t *out_K, int32_t *out_n)
{
    OraConst c; fold_constants(d, &c);
    v3 cam = { cam3[0], cam3[1], cam3[2] };
    const int TW = (d->W + 7) / 8, TH = (d->H + 7) / 8;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < TW * TH; ++tile) {
        int ti = tile / TH, tj = tile % TH;
        for (int ii = 0; ii < 8; ++ii) for (int jj = 0; jj < 8; ++jj) {
            int i = ti * 8 + ii, j = tj * 8 + jj;
            if (i >= d->W || j >= d->H) continue;
            size_t pix = (size_t)i * d->H + j;
            Ray r;
            setup_ray(d, &c, cam, i, j, jitter ? jitter[pix] : RC(0.0), &r);
            int K;
            v4 A = march(d, &c, vol, tf, &r, &K, NULL);
            if (d->nondiff) {                                          /* get_final_image_nondiff :358 */
                A.x = R_FMIN(RC(1.0), A.x); A.y = R_FMIN(RC(1.0), A.y); A.z = R_FMIN(RC(1.0), A.z); A.w = R_FMIN(RC(1.0), A.w);
            }
            out_rgba[4 * pix + 0] = A.x; out_rgba[4 * pix + 1] = A.y;
            out_rgba[4 * pix + 2] = A.z; out_rgba[4 * pix + 3] = A.w;
            if (out_K) out_K[pix] = K;
            if (out_n) out_n[pix] = r.n;
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Backward.  get_final_image.grad + raycast.grad for one item  :468-476, with Taichi's reverse-mode
 * rules restated.  The ray's A_{s-1} tape is kept per thread, exactly the
 * role render_tape plays in the reference (:300-302).
 *   grad_out [W][H][4] raw order
 *   gvol     [Y][Z][X] double, ACCUMULATED (caller zeroes)
 *   gtf      [R][4]    double, ACCUMULATED
 *   flags: bit0 = want volume grad, bit1 = want tf grad
 * Decisions for undefined corners: H4 (|g|==0 -> normal-path adjoint is 0), H8 (pow base clamp).
 * ------------------------------------------------------------------------------------------- */
void ora_backward(const OraDesc *d, const real *vol, const real *tf, const real *cam3,
                  const real *jitter, const real *grad_out, double *gvol, double *gtf, int flags)
{
    OraConst c; fold_constants(d, &c);
    v3 cam = { cam3[0], cam3[1], cam3[2] };
    const int want_vol = flags & 1, want_tf = flags & 2;
    const int TW = (d->W + 7) / 8, TH = (d->H + 7) / 8;
    const int R = d->R;
#pragma omp parallel
    {
        v4 *tape = NULL; int tape_cap = 0;
        double *ltf = (double *)calloc((size_t)R * 4, sizeof(double));
#pragma omp for schedule(dynamic, 4)
        for (int tile = 0; tile < TW * TH; ++tile) {
            int ti = tile / TH, tj = tile % TH;
            for (int ii = 0; ii < 8; ++ii) for (int jj = 0; jj < 8; ++jj) {
                int i = ti * 8 + ii, j = tj * 8 + jj;
                if (i >= d->W || j >= d->H) continue;
                size_t pix = (size_t)i * d->H + j;
                Ray r;
                setup_ray(d, &c, cam, i, j, jitter ? jitter[pix] : RC(0.0), &r);
                if (r.n <= 0) continue;
                if (r.n > tape_cap) { tape_cap = r.n + 64; tape = (v4 *)realloc(tape, sizeof(v4) * (size_t)tape_cap); }
                int K;
                (void)march(d, &c, vol, tf, &r, &K, tape);
                /* get_final_image.grad: tape.grad[ns-1] += out.grad ; inactive slots copy through */
                v4 g = { grad_out[4 * pix], grad_out[4 * pix + 1], grad_out[4 * pix + 2], grad_out[4 * pix + 3] };
                Sample q;
                for (int s = K - 1; s >= 0; --s) {
                    shade_sample(d, &c, vol, tf, &r, s, &q, 1);
                    v4 Ap = tape[s];
                    real T = RC(1.0) - Ap.w;
                    /* A_s = T*C + A_{s-1}  */
                    v4 dC = { T * g.x, T * g.y, T * g.z, T * g.w };
                    real Cg = q.C.x * g.x + q.C.y * g.y + q.C.z * g.z + q.C.w * g.w;
                    g.w = g.w - Cg;                                    /* d/dA_{s-1}.w through (1 - A.w) */
                    /* C = (k*c.rgb*o, o) */
                    real crgb_dC = q.c.x * dC.x + q.c.y * dC.y + q.c.z * dC.z;
                    real d_o = q.k * crgb_dC + dC.w;
                    real ko = q.k * q.o;
                    v4 dc = { ko * dC.x, ko * dC.y, ko * dC.z, RC(0.0) };
                    real dk = q.o * crgb_dC;
                    /* o = 1 - pow(1 - c.w, inv_sr) */
                    real base = RC(1.0) - q.c.w;
                    real dpow;
                    if (c.inv_sr == RC(1.0)) dpow = RC(1.0);
                    else dpow = c.inv_sr * R_POW(R_FMAX(base, RC(1e-12)), c.inv_sr - RC(1.0));   /* H8 */
                    dc.w = d_o * dpow;
                    /* tf lookup :215-219 */
                    const real *ta = tf + 4 * (size_t)q.lo, *tb = tf + 4 * (size_t)q.hi;
                    if (want_tf) {
                        real w0 = RC(1.0) - q.f, w1 = q.f;
                        ltf[4 * q.lo + 0] += (double)(dc.x * w0); ltf[4 * q.hi + 0] += (double)(dc.x * w1);
                        ltf[4 * q.lo + 1] += (double)(dc.y * w0); ltf[4 * q.hi + 1] += (double)(dc.y * w1);
                        ltf[4 * q.lo + 2] += (double)(dc.z * w0); ltf[4 * q.hi + 2] += (double)(dc.z * w1);
                        ltf[4 * q.lo + 3] += (double)(dc.w * w0); ltf[4 * q.hi + 3] += (double)(dc.w * w1);
                    }
                    if (!want_vol) continue;
                    real df = dc.x * (tb[0] - ta[0]) + dc.y * (tb[1] - ta[1]) + dc.z * (tb[2] - ta[2]) + dc.w * (tb[3] - ta[3]);
                    real dI = (q.x > RC(0.0)) ? df * c.tf_len : RC(0.0);     /* ti.max(x, 0) passes grad iff x > 0 */
                    trilinear_adjoint(&c, gvol, &q.cc, dI);
                    /* shading: k = min(1, .8*ndl + .3*rdv^32 + .4) */
                    if (q.glen2 > RC(0.0) && !(RC(1.0) < q.kraw)) {           /* H4 ; ti.min(1, k) passes grad iff !(1 < k) */
                        real d_ndl = (q.nl > RC(0.0)) ? RC(0.8) * dk : RC(0.0);
                        real d_rdv = (q.rv > RC(0.0)) ? RC(0.3) * RC(32.0) * R_POW(q.rdv, RC(31.0)) * dk : RC(0.0);
                        v3 dr = { -r.dir.x * d_rdv, -r.dir.y * d_rdv, -r.dir.z * d_rdv };
                        real drN = dot3(dr, q.N), drl = dot3(dr, q.l);
                        /* r = l - 2 (N.l) N ;  ndl = N.l  */
                        real cN = d_ndl - RC(2.0) * drN;   /* coefficient of l via d(N.l) */
                        (void)drl;
                        v3 dN = { cN * q.l.x - RC(2.0) * q.nl * dr.x,
                                  cN * q.l.y - RC(2.0) * q.nl * dr.y,
                                  cN * q.l.z - RC(2.0) * q.nl * dr.z };
                        real inv = RC(1.0) / R_SQRT(q.glen2);
                        real NdN = dot3(q.N, dN);
                        v3 dg = { (dN.x - q.N.x * NdN) * inv, (dN.y - q.N.y * NdN) * inv, (dN.z - q.N.z * NdN) * inv };
                        const real delta = RC(1e-3);
                        Cell t; v3 p = q.pos, a;
                        a = p; a.x = p.x + delta; locate(&c, a, &t); trilinear_adjoint(&c, gvol, &t, dg.x);
                        a = p; a.x = p.x - delta; locate(&c, a, &t); trilinear_adjoint(&c, gvol, &t, -dg.x);
                        a = p; a.y = p.y + delta; locate(&c, a, &t); trilinear_adjoint(&c, gvol, &t, dg.y);
                        a = p; a.y = p.y - delta; locate(&c, a, &t); trilinear_adjoint(&c, gvol, &t, -dg.y);
                        a = p; a.z = p.z + delta; locate(&c, a, &t); trilinear_adjoint(&c, gvol, &t, dg.z);
                        a = p; a.z = p.z - delta; locate(&c, a, &t); trilinear_adjoint(&c, gvol, &t, -dg.z);
                    }
                }
            }
        }
        if (want_tf) {
#pragma omp critical
            for (int q = 0; q < 4 * R; ++q) gtf[q] += ltf[q];
        }
        free(ltf); free(tape);
    }
}

int ora_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ora_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
