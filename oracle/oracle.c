/*
 * ORACLE -- test infrastructure, NOT product code.
 *
 * Plain-C Float64 restatement of the SemiLagrangian.jl hot path (the 1-D interpolation
 * sweep inside advection!).  The reference is Julia and cannot run here or on the GPU
 * box (no julia binary), so this file restates its algorithm function by function;
 * every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path never does.
 *
 * Parity status ("pinned" = checked against a known answer held by the reference's own
 * tests, see tests/test_oracle_*.py):
 *   - weight tables, LU layout, exact cubic-B-spline shift, analytic Float64 shifts,
 *     LU residuals, state schedule: PINNED by the reference's tests.
 *   - last-ulp behaviour (FMA contraction inside Base.evalpoly/muladd, @simd
 *     reassociation of sums with >= 16 terms, FFTW butterfly order): PARITY UNPINNED,
 *     the reference holds no golden vectors.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off).
 * -ffp-contract=off is deliberate: FMA is used only where the reference uses muladd.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXP 64 /* max order+1 */

/* ------------------------------------------------------------------------------------
 * src/interpolation.jl:96-98  getprecal(interp, decf) = [fct(decf) for fct in tabfct]
 * Polynomials.jl evaluates with Base.evalpoly == Horner with muladd (FMA on x86-64).
 * coef: np rows (one per stencil node) x nc ascending coefficients, row-major.
 * ---------------------------------------------------------------------------------- */
void orc_getprecal(const double *coef, int np, int nc, double t, double *w)
{
    for (int j = 0; j < np; ++j) {
        const double *c = coef + (size_t)j * nc;
        double ex = c[nc - 1];
        for (int k = nc - 2; k >= 0; --k)
            ex = fma(t, ex, c[k]);
        w[j] = ex;
    }
}

/* src/interpolation.jl:381-389  CachePrecal: decint = Int(floor(alpha)); decfloat = alpha - decint */
long orc_split_alpha(double alpha, double *decfloat)
{
    double fl = floor(alpha);
    *decfloat = alpha - fl;
    return (long)fl;
}

static inline long modn(long i, long n)
{
    long r = i % n;
    return r < 0 ? r + n : r;
}

/* ------------------------------------------------------------------------------------
 * src/interpolation.jl:175-193  interpolate!(fp, fi, decint, precal, interp::CircEdge)
 *   fp[i] = sum(res[tabmod[i+decal : i+decal+order]] .* precal)
 * products are rounded individually, then summed left to right (Base.sum is a
 * sequential loop for < 16 terms).  `res` = sol(interp, fi) is done by the caller.
 * ---------------------------------------------------------------------------------- */
void orc_interpolate_circ(double *fp, const double *res, long n, long decint, const double *w, int order)
{
    long origin = -(long)(order / 2);
    long decal = modn(origin + decint, n);
    for (long i = 0; i < n; ++i) {
        long k = i + decal;
        if (k >= n) k -= n;
        double s = 0.0;
        for (int j = 0; j <= order; ++j) {
            double prod = res[k] * w[j];
            s = (j == 0) ? prod : s + prod;
            if (++k == n) k = 0;
        }
        fp[i] = s;
    }
}

/* ------------------------------------------------------------------------------------
 * src/bsplinelu.jl:68-135  LuSpline(n, t; iscirc=true, isLU=true) + decLULu
 * 1-based accessors keep the transcription checkable against the Julia text.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    long n, szb;
    int wd, kl, ku;
    double *band;     /* wd x szb      */
    double *lastrows; /* ku x n        */
    double *lastcols; /* (n-ku) x kl   */
} orc_lu;

#define BAND(i, j) lu->band[((i)-1) + (size_t)((j)-1) * lu->wd]
#define LROW(i, j) lu->lastrows[((i)-1) + (size_t)((j)-1) * lu->ku]
#define LCOL(i, j) lu->lastcols[((i)-1) + (size_t)((j)-1) * (lu->n - lu->ku)]

static void orc_declulu(orc_lu *lu)
{ /* src/bsplinelu.jl:6-62, iscirc = true */
    int wd = lu->wd, kl = lu->kl, ku = lu->ku;
    long szb = lu->szb, n = lu->n;
    long begrow = n - ku, begcol = n - kl;
    for (long k = 1; k <= begrow; ++k) {
        double pivot = BAND(ku + 1, k);
        for (int i = ku + 2; i <= wd; ++i) BAND(i, k) /= pivot;
        for (int i = 1; i <= ku; ++i) LROW(i, k) /= pivot;
        for (int i = 1; i <= kl; ++i)
            for (int j = 1; j <= ku; ++j)
                if (k + j <= szb) BAND(ku + 1 + i - j, k + j) -= BAND(ku + 1 + i, k) * BAND(ku + 1 - j, k + j);
        long bkl = kl < begrow - k ? kl : begrow - k;
        for (int i = 1; i <= kl; ++i)
            for (long j = 1; j <= bkl; ++j) LCOL(k + j, i) -= BAND(ku + 1 + j, k) * LCOL(k, i);
        long bku = ku < szb - k ? ku : szb - k;
        for (long i = 1; i <= bku; ++i)
            for (int j = 1; j <= ku; ++j) LROW(j, k + i) -= LROW(j, k) * BAND(ku + 1 - i, k + i);
        for (int i = 1; i <= ku; ++i)
            for (int j = 1; j <= kl; ++j) LROW(i, begcol + j) -= LROW(i, k) * LCOL(k, j);
    }
    for (long k = begrow + 1; k <= n; ++k) {
        long i_k = k - begrow;
        double pivot = LROW(i_k, k);
        for (long i = i_k + 1; i <= ku; ++i) LROW(i, k) /= pivot;
        for (long i = i_k + 1; i <= ku; ++i)
            for (long j = k + 1; j <= n; ++j) LROW(i, j) -= LROW(i, k) * LROW(k - begrow, j);
    }
}

void orc_lu_destroy(orc_lu *lu)
{
    if (!lu) return;
    free(lu->band);
    free(lu->lastrows);
    free(lu->lastcols);
    free(lu);
}

/* t = B(1..order) (wd = order values); src/bsplinelu.jl:89-135 */
orc_lu *orc_lu_create(long n, const double *t, int wd, int do_lu)
{
    orc_lu *lu = (orc_lu *)calloc(1, sizeof(orc_lu));
    int ku = wd / 2, kl = wd - 1 - ku; /* src/bspline.jl:12-16 get_kl_ku(wd) */
    lu->n = n; lu->wd = wd; lu->kl = kl; lu->ku = ku;
    long szb = n - kl;
    lu->szb = szb;
    lu->band = (double *)calloc((size_t)wd * szb, sizeof(double));
    lu->lastrows = (double *)calloc((size_t)ku * n + 1, sizeof(double));
    lu->lastcols = (double *)calloc((size_t)(n - ku) * kl + 1, sizeof(double));
    for (int i = 1; i <= wd; ++i) {
        long jbeg = i <= ku + 1 ? ku - i + 2 : 1;
        long jend = i >= kl + 2 ? szb - i + kl + 1 : szb;
        for (long j = jbeg; j <= jend; ++j) BAND(i, j) = t[i - 1];
    }
    for (int i = 1; i <= ku; ++i)
        for (int ind = 1; ind <= wd; ++ind) {
            long j = n - wd + i + ind;
            LROW(i, (j - 1) % n + 1) = t[ind - 1];
        }
    for (int i = 1; i <= kl; ++i)
        for (int j = 1; j <= kl + 1 - i; ++j) {
            LCOL(j, j + i - 1) = t[i - 1];
            LCOL(n - kl - ku + i + j - 1, j) = t[i - 1];
        }
    if (do_lu) orc_declulu(lu);
    return lu;
}

/* accessors for the layout test (tests/test_oracle_lu.py) */
long orc_lu_dims(const orc_lu *lu, long *out /* n, szb, wd, kl, ku */)
{
    out[0] = lu->n; out[1] = lu->szb; out[2] = lu->wd; out[3] = lu->kl; out[4] = lu->ku;
    return 0;
}
const double *orc_lu_band(const orc_lu *lu) { return lu->band; }
const double *orc_lu_lastrows(const orc_lu *lu) { return lu->lastrows; }
const double *orc_lu_lastcols(const orc_lu *lu) { return lu->lastcols; }

/* ------------------------------------------------------------------------------------
 * src/bsplinelu.jl:179-220  sol!(X, spA::LuSpline, Y)   (Y is overwritten, X = solution)
 * sums are sequential left-to-right (Julia: < 16 terms, or generator => foldl);
 * the dense lastrows dot products (:194) have ~n terms and are @simd-reassociable in
 * Julia -- sequential here (parity unpinned at the ulp level).
 * ---------------------------------------------------------------------------------- */
void orc_lu_sol(const orc_lu *lu, double *X, double *Y)
{
    long n = lu->n;
    int kl = lu->kl, ku = lu->ku;
    long begrow = n - ku, begcol = n - kl;
    long endmat = begrow, endmat2 = begcol;
#define y(i) Y[(i)-1]
#define x(i) X[(i)-1]
    for (long i = 2; i <= endmat; ++i) {
        long fin = i - 1, deb = i - kl > 1 ? i - kl : 1;
        double s = 0.0;
        for (long j = deb; j <= fin; ++j) {
            double p = y(j) * BAND(ku + 1 + i - j, j);
            s = (j == deb) ? p : s + p;
        }
        y(i) -= s;
    }
    for (long i = begrow + 1; i <= n; ++i) {
        double s = 0.0;
        for (long j = 1; j <= i - 1; ++j) {
            double p = y(j) * LROW(i - begrow, j);
            s = (j == 1) ? p : s + p;
        }
        y(i) -= s;
    }
    for (long i = 1; i <= n; ++i) x(i) = 0.0;
    for (long i = n; i >= begrow + 1; --i) {
        double s = 0.0;
        int first = 1;
        for (long j = i + 1; j <= n; ++j) {
            double p = x(j) * LROW(i - begrow, j);
            s = first ? p : s + p;
            first = 0;
        }
        x(i) = (y(i) - s) / LROW(i - begrow, i);
    }
    for (long i = endmat; i >= 1; --i) {
        long deb = i + 1, fin = i + ku < endmat2 ? i + ku : endmat2;
        double s = 0.0;
        if (deb <= fin) {
            for (long j = deb; j <= fin; ++j) {
                double p = BAND(ku + 1 + i - j, j) * x(j);
                s = (j == deb) ? p : s + p;
            }
        }
        double s2 = 0.0;
        for (int j = 1; j <= kl; ++j) {
            double p = LCOL(i, j) * x(n - kl + j);
            s2 = (j == 1) ? p : s2 + p;
        }
        s += s2;
        x(i) = (y(i) - s) / BAND(ku + 1, i);
    }
#undef x
#undef y
}

/* ------------------------------------------------------------------------------------
 * Float64 DFT convention of src/fftbig.jl:162-176,194-208 (FFTW): forward
 * exp(-2 pi i jk/n) unnormalised, inverse exp(+...)/n.  Iterative radix-2 (the
 * reference asserts n is a power of two, src/fftbig.jl:57).  FFTW's butterfly order is
 * not reproduced: parity unpinned at the ulp level.
 * ---------------------------------------------------------------------------------- */
static void orc_fft_pow2(double *re, double *im, long n, int inverse)
{
    for (long i = 1, j = 0; i < n; ++i) {
        long bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    const double PI2 = 6.283185307179586476925286766559;
    for (long len = 2; len <= n; len <<= 1) {
        long half = len >> 1;
        for (long i = 0; i < n; i += len)
            for (long k = 0; k < half; ++k) {
                double ang = (inverse ? PI2 : -PI2) * (double)k / (double)len;
                double wr = cos(ang), wi = sin(ang);
                double ur = re[i + k], ui = im[i + k];
                double vr = re[i + k + half] * wr - im[i + k + half] * wi;
                double vi = re[i + k + half] * wi + im[i + k + half] * wr;
                re[i + k] = ur + vr; im[i + k] = ui + vi;
                re[i + k + half] = ur - vr; im[i + k + half] = ui - vi;
            }
    }
    if (inverse)
        for (long i = 0; i < n; ++i) { re[i] /= (double)n; im[i] /= (double)n; }
}

int orc_fft(double *re, double *im, long n, int inverse)
{
    if (n <= 0 || (n & (n - 1))) return -1;
    orc_fft_pow2(re, im, n, inverse);
    return 0;
}

/* src/bsplinefft.jl:29-42: c[(n-kl-1+i) % n + 1] = B(i), i=1..order; c_fft = fft(c) */
typedef struct {
    long n;
    double *cre, *cim;
} orc_bfft;

void orc_bfft_destroy(orc_bfft *b)
{
    if (!b) return;
    free(b->cre); free(b->cim); free(b);
}

orc_bfft *orc_bfft_create(long n, const double *t, int order)
{
    if (n <= 0 || (n & (n - 1))) return NULL; /* src/fftbig.jl:57 */
    orc_bfft *b = (orc_bfft *)calloc(1, sizeof(orc_bfft));
    b->n = n;
    b->cre = (double *)calloc(n, sizeof(double));
    b->cim = (double *)calloc(n, sizeof(double));
    int ku = order / 2, kl = order - 1 - ku;
    long dec = n - kl - 1;
    for (int i = 1; i <= order; ++i) b->cre[(dec + i) % n] = t[i - 1];
    orc_fft_pow2(b->cre, b->cim, n, 0);
    return b;
}

/* src/bsplinefft.jl:49-51: real(ifft(fft(b) ./ c_fft)) */
void orc_bfft_sol(const orc_bfft *bf, double *X, const double *b, double *wre, double *wim)
{
    long n = bf->n;
    for (long i = 0; i < n; ++i) { wre[i] = b[i]; wim[i] = 0.0; }
    orc_fft_pow2(wre, wim, n, 0);
    for (long i = 0; i < n; ++i) {
        double a = wre[i], bb = wim[i], c = bf->cre[i], d = bf->cim[i];
        double den = c * c + d * d;
        wre[i] = (a * c + bb * d) / den;
        wim[i] = (bb * c - a * d) / den;
    }
    orc_fft_pow2(wre, wim, n, 1);
    for (long i = 0; i < n; ++i) X[i] = wre[i];
}

/* ------------------------------------------------------------------------------------
 * interpolation descriptor shared by the line and sweep entry points
 * kind: 0 Lagrange, 1 BSplineLU, 2 BSplineFFT, 3 Hermite
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int kind, order, nc;
    long n;
    double *coef; /* (order+1) x nc */
    orc_lu *lu;
    orc_bfft *bf;
} orc_interp;

void orc_interp_destroy(orc_interp *it)
{
    if (!it) return;
    free(it->coef);
    orc_lu_destroy(it->lu);
    orc_bfft_destroy(it->bf);
    free(it);
}

orc_interp *orc_interp_create(int kind, int order, long n, const double *coef, int nc, const double *node_vals)
{
    if (order + 1 > ORC_MAXP) return NULL;
    orc_interp *it = (orc_interp *)calloc(1, sizeof(orc_interp));
    it->kind = kind; it->order = order; it->nc = nc; it->n = n;
    it->coef = (double *)malloc(sizeof(double) * (order + 1) * nc);
    memcpy(it->coef, coef, sizeof(double) * (order + 1) * nc);
    if (kind == 1) it->lu = orc_lu_create(n, node_vals, order, 1);
    if (kind == 2) {
        it->bf = orc_bfft_create(n, node_vals, order);
        if (!it->bf) { orc_interp_destroy(it); return NULL; }
    }
    return it;
}

/* sol(interp, fi): identity (src/interpolation.jl:40), LU (src/bsplinelu.jl:282-284), FFT (src/bsplinefft.jl:49-51).
 * work must hold 3n doubles. Returns pointer to the coefficients (either fi or work). */
static const double *orc_sol(const orc_interp *it, const double *fi, double *work)
{
    long n = it->n;
    if (it->kind == 1) {
        double *X = work, *Y = work + n;
        memcpy(Y, fi, sizeof(double) * n);
        orc_lu_sol(it->lu, X, Y);
        return X;
    }
    if (it->kind == 2) {
        orc_bfft_sol(it->bf, work, fi, work + n, work + 2 * n);
        return work;
    }
    return fi;
}

/* One line, full reference semantics of the const-shift branch of advection!
 * (src/advection.jl:627-631): getprecal(cache, alpha) -> interpolate! -> slc .= buf.
 * fp and fi must not alias. */
void orc_interpolate_alpha(const orc_interp *it, double *fp, const double *fi, long n, double alpha)
{
    double w[ORC_MAXP], t;
    double *work = (double *)malloc(sizeof(double) * 3 * n);
    long decint = orc_split_alpha(alpha, &t);
    orc_getprecal(it->coef, it->order + 1, it->nc, t, w);
    const double *res = orc_sol(it, fi, work);
    orc_interpolate_circ(fp, res, n, decint, w, it->order);
    free(work);
}

/* just the pre-solve (for tests of sol) */
void orc_sol_line(const orc_interp *it, double *X, const double *fi)
{
    long n = it->n;
    double *work = (double *)malloc(sizeof(double) * 3 * n);
    const double *r = orc_sol(it, fi, work);
    memcpy(X, r, sizeof(double) * n);
    free(work);
}

/* ------------------------------------------------------------------------------------
 * InsideEdge (non-periodic) interpolation, "a marginal case" in the reference's words:
 * src/interpolation.jl:123-132  get_allprecal: allprecal[k] = getprecal(interp, decfloat + indbeg + k - 1),
 *                                k = 1..order+1, indbeg = -div(order, 2) + decint
 * src/interpolation.jl:250-286  interpolate!(fp, fi, decint, allprecal, interp::InsideEdge):
 *   the window never leaves the array: the first borne1 = -indbeg outputs use res[1:order+1], the last
 *   ones res[lg-order:lg], each with the weights of its own offset; in between the standard window.
 * Products rounded, summed left to right.  Returns -1 when the reference's loops would index outside
 * the arrays (indbeg > 0 or indbeg + order < 0).
 * ---------------------------------------------------------------------------------- */
int orc_interpolate_inside(double *fp, const double *res, long lg, long decint, double decfloat, const double *coef,
                           int order, int nc)
{
    long origin = -(long)(order / 2);
    long indbeg = origin + decint;
    long borne1 = -decint - origin;           /* = -indbeg */
    long borne2 = lg - decint + origin - 1;
    int lgp = order + 1;
    if (borne1 < 0 || borne1 > order || lg < order + 1) return -1;
    double w[ORC_MAXP];
    for (long i = 1; i <= lg; ++i) {          /* 1-based, as the Julia text */
        long first, ind;
        if (i <= borne1) { first = 1; ind = i; }
        else if (i <= borne2) { first = i - borne1; ind = borne1 + 1; }
        else { first = lg - order; ind = lgp - (lg - i); }
        orc_getprecal(coef, order + 1, nc, decfloat + (double)(indbeg + ind - 1), w);
        double s = 0.0;
        for (int j = 0; j <= order; ++j) {
            double prod = res[first - 1 + j] * w[j];
            s = (j == 0) ? prod : s + prod;
        }
        fp[i - 1] = s;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * src/interpolation.jl:561-621 (and its closure twin :401-429), N = 2:
 * interpolate!(fp, fi, bufdec::Array{OpTuple{2}}, interp_t).  `res` = sol(interp_t, fi) is
 * done by the caller (per dim, src/interpolation.jl:48-94).  Per point ind = (i, j):
 *     dint, tab = getprecal(cache, bufdec[ind])          :345-353: dint = Int.(floor.(alpha)),
 *                                                         tab = dotprod(getprecal per dim) (:112-118)
 *     deb_i = dint .+ decall .+ ind.I, decall = (5sz + origin) % sz + sz, origin = -div(order, 2)
 *     fp[ind] = sum(res[tabmod[1][deb:end], tabmod[2][deb:end]] .* tab)
 * tab[a, b] = wA[a] * wB[b] (rounded); the products res .* tab are rounded, then summed in
 * column-major order (a fastest).  Julia's sum of a (p+1)^2 array is a plain loop below 1024
 * elements; @simd reassociation is unpinned (see the header).
 * Arrays: res, fp [n1, n2, ncomp] column-major (ncomp planes = the components of an OpTuple
 * field), dec [n1, n2, 2].  fp must not alias res.
 * ---------------------------------------------------------------------------------- */
void orc_interpolate_points2d(const orc_interp *itA, const orc_interp *itB, double *fp, const double *res,
                              const double *dec, long n1, long n2, int ncomp, int nthreads)
{
    const int pA = itA->order + 1, pB = itB->order + 1;
    const long plane = n1 * n2;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (long j = 0; j < n2; ++j) {
        double wA[ORC_MAXP], wB[ORC_MAXP], tA, tB;
        double tab[ORC_MAXP * ORC_MAXP];
        for (long i = 0; i < n1; ++i) {
            long idx = i + n1 * j;
            long dA = orc_split_alpha(dec[idx], &tA);
            long dB = orc_split_alpha(dec[plane + idx], &tB);
            orc_getprecal(itA->coef, pA, itA->nc, tA, wA);
            orc_getprecal(itB->coef, pB, itB->nc, tB, wB);
            for (int b = 0; b < pB; ++b)
                for (int a = 0; a < pA; ++a) tab[a + pA * b] = wA[a] * wB[b];
            long ia0 = modn(i + dA - itA->order / 2, n1);
            long jb0 = modn(j + dB - itB->order / 2, n2);
            for (int c = 0; c < ncomp; ++c) {
                const double *r = res + (size_t)c * plane;
                double s = 0.0;
                long jb = jb0;
                for (int b = 0; b < pB; ++b) {
                    long ia = ia0;
                    for (int a = 0; a < pA; ++a) {
                        double prod = r[ia + n1 * jb] * tab[a + pA * b];
                        s = (a == 0 && b == 0) ? prod : s + prod;
                        if (++ia == n1) ia = 0;
                    }
                    if (++jb == n2) jb = 0;
                }
                fp[(size_t)c * plane + idx] = s;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * permutedims! with the advected dim first (src/advection.jl:372-376) and back
 * (:385).  data is column-major with extents ext[0..nd); view = [inner, n, outer].
 *   f[k + n*(a + inner*b)] = data[a + inner*(k + n*b)]
 * ---------------------------------------------------------------------------------- */
static void permute_fwd(double *f, const double *data, long inner, long n, long outer)
{
    const long B = 32;
#pragma omp parallel for collapse(2) schedule(static)
    for (long b = 0; b < outer; ++b)
        for (long a0 = 0; a0 < inner; a0 += B) {
            const double *src = data + (size_t)inner * n * b;
            double *dst = f + (size_t)inner * n * b;
            long a1 = a0 + B < inner ? a0 + B : inner;
            for (long k0 = 0; k0 < n; k0 += B) {
                long k1 = k0 + B < n ? k0 + B : n;
                for (long a = a0; a < a1; ++a)
                    for (long k = k0; k < k1; ++k) dst[k + n * a] = src[a + inner * k];
            }
        }
}

static void permute_bwd(double *data, const double *f, long inner, long n, long outer)
{
    const long B = 32;
#pragma omp parallel for collapse(2) schedule(static)
    for (long b = 0; b < outer; ++b)
        for (long a0 = 0; a0 < inner; a0 += B) {
            double *dst = data + (size_t)inner * n * b;
            const double *src = f + (size_t)inner * n * b;
            long a1 = a0 + B < inner ? a0 + B : inner;
            for (long k0 = 0; k0 < n; k0 += B) {
                long k1 = k0 + B < n ? k0 + B : n;
                for (long k = k0; k < k1; ++k)
                    for (long a = a0; a < a1; ++a) dst[a + inner * k] = src[k + n * a];
            }
        }
}

/* ------------------------------------------------------------------------------------
 * One advection! call for a const-shift 1-D state (src/advection.jl:594-657):
 *   f = getformdata (permute), for every trailing index: getprecal(cache, getalpha)
 *   -> interpolate!(buf, slc, ...) -> slc .= buf, then copydata! (permute back).
 * alpha for the line whose other-dim indices are idx[] is
 *   alpha_tab[ sum_d idx[d]*astride[d] ]   (astride[dim] ignored)
 * which is how the plugins' bufcur tables are indexed (src/poisson.jl:210-224,
 * src/rotation.jl:71, src/translation.jl:33-35).
 * nthreads <= 1: NoTimeOpt (:622-632); > 1: SimpleThreadsOpt (:647-657, @threads over
 * lines with per-thread buffer and per-thread CachePrecal).
 * scratch must hold numel doubles (AdvectionData.bufdata, src/advection.jl:265).
 * ---------------------------------------------------------------------------------- */
int orc_sweep(double *data, double *scratch, int nd, const long *ext, int dim, const orc_interp *it,
              const double *alpha_tab, const long *astride, int nthreads)
{
    if (dim < 0 || dim >= nd) return -1;
    long n = ext[dim];
    if (n != it->n && it->kind != 0 && it->kind != 3) return -2;
    long inner = 1, outer = 1;
    for (int d = 0; d < dim; ++d) inner *= ext[d];
    for (int d = dim + 1; d < nd; ++d) outer *= ext[d];
    int order = it->order, np = order + 1;
#ifdef _OPENMP
    int saved = omp_get_max_threads();
    omp_set_num_threads(nthreads > 1 ? nthreads : 1);
#endif
    permute_fwd(scratch, data, inner, n, outer);
    long nlines = inner * outer;
#pragma omp parallel
    {
        double *buf = (double *)malloc(sizeof(double) * n);
        double *work = (double *)malloc(sizeof(double) * 3 * n);
        double w[ORC_MAXP];
        /* CachePrecal (src/interpolation.jl:320-349): initialised with alpha = 0 */
        double cache_alpha = 0.0, t0 = 0.0;
        long cache_int = 0;
        orc_getprecal(it->coef, np, it->nc, t0, w);
#pragma omp for schedule(static)
        for (long line = 0; line < nlines; ++line) {
            long a = line % inner, b = line / inner;
            long off = 0, r = a;
            for (int d = 0; d < dim; ++d) { off += (r % ext[d]) * astride[d]; r /= ext[d]; }
            r = b;
            for (int d = dim + 1; d < nd; ++d) { off += (r % ext[d]) * astride[d]; r /= ext[d]; }
            double alpha = alpha_tab[off];
            if (alpha != cache_alpha) { /* src/interpolation.jl:381-389 */
                cache_alpha = alpha;
                double t;
                cache_int = orc_split_alpha(alpha, &t);
                orc_getprecal(it->coef, np, it->nc, t, w);
            }
            double *slc = scratch + (size_t)line * n;
            const double *res = orc_sol(it, slc, work);
            orc_interpolate_circ(buf, res, n, cache_int, w, order);
            memcpy(slc, buf, sizeof(double) * n);
        }
        free(buf);
        free(work);
    }
    permute_bwd(data, scratch, inner, n, outer);
#ifdef _OPENMP
    omp_set_num_threads(saved);
#endif
    return 0;
}

/* ------------------------------------------------------------------------------------
 * src/util_poisson.jl:68-79 compute_charge!:  rho = dv * sum_v f ; rho -= mean(rho)
 * f viewed as [nsp, nv] column-major; Julia's sum(f; dims) accumulates r[a] += f[a,b]
 * in memory order, i.e. sequentially over b for each a.  The mean uses a plain
 * sequential sum here (Julia: pairwise blocks of 1024 -- ulp-level, unpinned).
 * ---------------------------------------------------------------------------------- */
void orc_compute_charge(double *rho, const double *f, long nsp, long nv, double dv, int nthreads)
{
#ifdef _OPENMP
    int saved = omp_get_max_threads();
    omp_set_num_threads(nthreads > 1 ? nthreads : 1);
#endif
#pragma omp parallel for schedule(static)
    for (long a0 = 0; a0 < nsp; a0 += 512) {
        long a1 = a0 + 512 < nsp ? a0 + 512 : nsp;
        for (long a = a0; a < a1; ++a) rho[a] = 0.0;
        for (long b = 0; b < nv; ++b) {
            const double *p = f + (size_t)nsp * b;
            for (long a = a0; a < a1; ++a) rho[a] += p[a];
        }
        for (long a = a0; a < a1; ++a) rho[a] = dv * rho[a];
    }
#ifdef _OPENMP
    omp_set_num_threads(saved);
#endif
    double s = 0.0;
    for (long a = 0; a < nsp; ++a) s += rho[a];
    double mean = s / (double)nsp;
    for (long a = 0; a < nsp; ++a) rho[a] -= mean;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
