/* oracle.c - TEST INFRASTRUCTURE ONLY.  Never linked, imported or called by the product
 * (mcsolver_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and there only as the checker / the timed CPU baseline.
 *
 * A plain-C, array-based CPU restatement of the reference engines of golddoushi/mcsolver
 *   heisenbergLib.c (O(3)), xyLib.c (O(2)), isingLib.c (O(1)).
 * Each function cites the reference file:line it follows.  Two kinds of entry points:
 *
 *  (1) FAITHFUL restatement (update modes 0/1): same arithmetic expression order and the same
 *      libc rand() call sequence as the reference, so that after srand(k) orc_run()/
 *      orc_run_ising() reproduce the reference's MCMainFunction result tuple.  This is how
 *      the oracle is PINNED: tests/test_oracle_pin.py compares it with the reference's own
 *      compiled code (oracle/_ref, built by oracle/Makefile from /root/reference) and with
 *      the committed golden fixtures generated from it (tests/golden/).
 *      Known reference defects are reproduced behind flags (isingStrideBug, wolffHalfMove) so
 *      the pin is exact; the CUDA engine implements the intended physics (flags = 0).
 *
 *  (2) SCHEDULING variants (update modes 2/3): the same per-attempt physics functions
 *      (delta energy, acceptance rule, cluster rule), but sites are visited in colour-class
 *      order with a counter-based Philox4x32-10 stream instead of random sites with rand().
 *      This is the algorithm the CUDA engine runs; it lets the GPU path be checked
 *      deterministically (same trajectories in fp64) rather than only statistically.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define REF_PI 3.1415926535 /* heisenbergLib.c:6 (truncated on purpose: Q is off-integer by ~3e-11) */

typedef struct {
    int model;            /* 1 Ising, 2 XY, 3 Heisenberg */
    int N, maxL;
    const double *S;      /* [N] signed initial spin / spin length */
    const double *D;      /* [N*3] single-ion anisotropy (already /T); NULL for Ising */
    const int *nlink;     /* [N] */
    const double *J;      /* O(n): [N*maxL*9] xx,yy,zz,xy,xz,yz,yx,zx,zy ; Ising: [N*maxL] */
    const int *nbr;       /* [N*maxL], -1 padded */
    int nTri;
    const int *tri;       /* [nTri*3] */
    int nLat;
    const int *pairs;     /* [nLat*2] */
    int nG, maxG;
    const int *groups;    /* [nG*maxG], -1 padded */
    int nR, nC;
    const int *rOrb;      /* [nR] */
    const int *rCl;       /* [nR*nC] */
    const int *rNbr;      /* [nR*maxL] */
    double h;             /* H/T */
    int ignoreOffDiag;    /* p_diagonalDot = diagonalDot_simple (heisenbergLib.c:592-593) */
    int isingStrideBug;   /* reproduce isingLib.c:30 (linkStrength+i instead of +i*maxNLinking) */
    int wolffHalfMove;    /* reproduce heisenbergLib.c:418 / xyLib.c:362 (residual uses half move) */
    int rngStride;        /* Philox colour sweeps: 0 = one Philox block per site (table path); S >= 1 = grouped streams of the */
    int rngGroup;         /*   vectorised structured pass: rngGroup sites with ids base + m*S share their blocks (csrc/rng.cuh) */
} orc_sys;

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11) - shared convention with the CUDA engine (rng.cuh)     */
/* ------------------------------------------------------------------------------------------ */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* counter layout: (site, sweep_lo, sweep_hi16 | sub<<16 | stream<<24, replica); key = seed */
enum { STREAM_METRO = 0, STREAM_INIT = 1, STREAM_WBOND = 2, STREAM_WSEED = 3, STREAM_PT = 4 };
static void rng4(uint64_t seed, uint32_t replica, uint32_t stream, uint32_t sub, uint64_t sweep,
                 uint32_t site, uint32_t out[4]) {
    uint32_t ctr[4] = {site, (uint32_t)sweep, (uint32_t)((sweep >> 32) & 0xFFFFu) | (sub << 16) | (stream << 24), replica};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32(ctr, key, out);
}
/* Words of site id for one colour sweep, in the slots the attempt code reads: r[0], r[1] direction, r[2] acceptance,
 * r[3] attempt probability (partial sweeps).  W = words per attempt = number of spin components.
 *   stride 0 : per-site blocks (counter word 0 = id); Ising: four consecutive ids share the block id >> 2, word id & 3,
 *              sub-stream 1 for the attempt-probability uniform.
 *   stride S : grouped streams - group G = (id/(V*S))*S + id%S, member m = (id/S)%V, word k = W*m + t is word k&3 of the
 *              block drawn with sub-stream k>>2; attempt probability: word m&3 of sub-stream 7+(m>>2).  (csrc/rng.cuh) */
static void site_words(const orc_sys *s, uint64_t seed, uint32_t replica, uint64_t sweep, uint32_t id, int partial, uint32_t r[4]) {
    const int W = s->model;
    uint32_t b[4];
    r[0] = r[1] = r[2] = r[3] = 0;
    if (s->rngStride <= 0) {
        if (W == 1) {
            rng4(seed, replica, STREAM_METRO, 0, sweep, id >> 2, b);
            r[2] = b[id & 3];
            if (partial) { rng4(seed, replica, STREAM_METRO, 1, sweep, id >> 2, b); r[3] = b[id & 3]; }
        } else rng4(seed, replica, STREAM_METRO, 0, sweep, id, r);
        return;
    }
    const uint32_t S = (uint32_t)s->rngStride, V = (uint32_t)s->rngGroup;
    const uint32_t G = (id / (V * S)) * S + id % S, m = (id / S) % V;
    uint32_t w[3] = {0, 0, 0};
    for (int t = 0; t < W; t++) {
        uint32_t k = (uint32_t)W * m + (uint32_t)t;
        rng4(seed, replica, STREAM_METRO, k >> 2, sweep, G, b);
        w[t] = b[k & 3];
    }
    if (W == 3) { r[0] = w[0]; r[1] = w[1]; r[2] = w[2]; }
    else if (W == 2) { r[0] = w[0]; r[2] = w[1]; }
    else r[2] = w[0];
    if (partial) { rng4(seed, replica, STREAM_METRO, 7 + (m >> 2), sweep, G, b); r[3] = b[m & 3]; }
}
static double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }
/* fp32 engine convention: 23 random bits, (k+0.5)/2^23, evaluated exactly in double here */
static double u01f(uint32_t r) { return ((double)(r >> 9) + 0.5) * (1.0 / 8388608.0); }

/* ------------------------------------------------------------------------------------------ */
/* physics of one configuration (spins: O(n) [N*3] with unused comps 0; Ising [N])            */
/* ------------------------------------------------------------------------------------------ */
/* diagonalDot / diagonalDot_simple - heisenbergLib.c:77-93, xyLib.c:59-68 */
static double ddot(const orc_sys *s, const double *v1, const double *v2, const double *J) {
    if (s->model == 3) {
        if (s->ignoreOffDiag) return v1[0] * v2[0] * J[0] + v1[1] * v2[1] * J[1] + v1[2] * v2[2] * J[2];
        return v1[0] * v2[0] * J[0] + v1[1] * v2[1] * J[1] + v1[2] * v2[2] * J[2] + v1[0] * v2[1] * J[3] +
               v1[0] * v2[2] * J[4] + v1[1] * v2[2] * J[5] + v1[1] * v2[0] * J[6] + v1[2] * v2[0] * J[7] +
               v1[2] * v2[1] * J[8];
    }
    /* XY takes flat idx 0,1,3,6 = xx,yy,xy,yx (xyLib.c:150-153) */
    if (s->ignoreOffDiag) return v1[0] * v2[0] * J[0] + v1[1] * v2[1] * J[1];
    return v1[0] * v2[0] * J[0] + v1[1] * v2[1] * J[1] + v1[0] * v2[1] * J[3] + v1[1] * v2[0] * J[6];
}
static double vdot(const orc_sys *s, const double *a, const double *b) {
    if (s->model == 3) return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    return a[0] * b[0] + a[1] * b[1];
}
static const double *isingJ(const orc_sys *s, int i) {
    return s->isingStrideBug ? s->J + i : s->J + (size_t)i * s->maxL; /* isingLib.c:30 */
}
/* getCorrEnergy - heisenbergLib.c:238-247, xyLib.c:191-198, isingLib.c:121-127 */
double orc_site_bond_energy(const orc_sys *s, const double *sp, int i) {
    double corr = 0;
    if (s->model == 1) {
        const double *J = isingJ(s, i);
        for (int k = 0; k < s->nlink[i]; k++) corr += J[k] * sp[i] * sp[s->nbr[(size_t)i * s->maxL + k]];
        return corr;
    }
    for (int k = 0; k < s->nlink[i]; k++) {
        int j = s->nbr[(size_t)i * s->maxL + k];
        corr += ddot(s, sp + 3 * i, sp + 3 * j, s->J + ((size_t)i * s->maxL + k) * 9);
    }
    return corr;
}
/* getOnsiteEnergy - heisenbergLib.c:249-253 (field along z), xyLib.c:200-204 (field along x) */
double orc_site_onsite_energy(const orc_sys *s, const double *sp, int i) {
    if (s->model == 1) return -s->h * sp[i]; /* isingLib.c:232 */
    const double *v = sp + 3 * i, *D = s->D + 3 * i;
    if (s->model == 3) return D[0] * v[0] * v[0] + D[1] * v[1] * v[1] + D[2] * v[2] * v[2] - s->h * v[2];
    return D[0] * v[0] * v[0] + D[1] * v[1] * v[1] - s->h * v[0];
}
/* total energy - heisenbergLib.c:602-606 / 432-435; Ising absolute form isingLib.c:230-233 */
double orc_total_energy(const orc_sys *s, const double *sp) {
    double e = 0;
    if (s->model == 1) {
        for (int i = 0; i < s->N; i++) e += orc_site_bond_energy(s, sp, i) / 2 - s->h * sp[i];
        return e;
    }
    for (int i = 0; i < s->N; i++) e += orc_site_bond_energy(s, sp, i);
    e /= 2;
    for (int i = 0; i < s->N; i++) e += orc_site_onsite_energy(s, sp, i);
    return e;
}
/* per-site energies (bond part un-halved) for the engine's mcg_energy hook */
void orc_site_energies(const orc_sys *s, const double *sp, double *ebond, double *eons) {
    for (int i = 0; i < s->N; i++) {
        ebond[i] = orc_site_bond_energy(s, sp, i);
        eons[i] = orc_site_onsite_energy(s, sp, i);
    }
}
/* getDeltaCorrEnergy + getDeltaOnsiteEnergy - heisenbergLib.c:288-308, xyLib.c:239-254 */
double orc_delta_energy(const orc_sys *s, const double *sp, int i, const double *trans) {
    double corr = 0;
    for (int k = 0; k < s->nlink[i]; k++) {
        int j = s->nbr[(size_t)i * s->maxL + k];
        corr += ddot(s, trans, sp + 3 * j, s->J + ((size_t)i * s->maxL + k) * 9);
    }
    const double *v = sp + 3 * i, *D = s->D + 3 * i;
    double s1x = v[0] + trans[0], s1y = v[1] + trans[1], s1z = v[2] + trans[2];
    double on;
    if (s->model == 3)
        on = D[0] * (s1x * s1x - v[0] * v[0]) + D[1] * (s1y * s1y - v[1] * v[1]) + D[2] * (s1z * s1z - v[2] * v[2]) -
             s->h * trans[2];
    else
        on = D[0] * (s1x * s1x - v[0] * v[0]) + D[1] * (s1y * s1y - v[1] * v[1]) - s->h * trans[0];
    return corr + on;
}
/* Ising flip "corr" - isingLib.c:242: 2*(sum J s_i s_j - h s_i); flip lowers E by corr */
double orc_ising_flip_corr(const orc_sys *s, const double *sp, int i) {
    return 2 * (orc_site_bond_energy(s, sp, i) - s->h * sp[i]);
}
/* calcSignedArea - heisenbergLib.c:114-127 */
double orc_signed_area(const double *s1, const double *s2, const double *s3, double l1, double l2, double l3) {
    double s1s2 = (s1[0] * s2[0] + s1[1] * s2[1] + s1[2] * s2[2]) / l1 / l2;
    double s2s3 = (s2[0] * s3[0] + s2[1] * s3[1] + s2[2] * s3[2]) / l2 / l3;
    double s3s1 = (s3[0] * s1[0] + s3[1] * s1[1] + s3[2] * s1[2]) / l3 / l1;
    double cx = s2[1] * s3[2] - s2[2] * s3[1], cy = s2[2] * s3[0] - s2[0] * s3[2], cz = s2[0] * s3[1] - s2[1] * s3[0];
    double re = 1 + s1s2 + s2s3 + s3s1;
    double im = (s1[0] * cx + s1[1] * cy + s1[2] * cz) / l1 / l2 / l3;
    if (fabs(re) < 1e-6) return im > 0 ? REF_PI : -REF_PI;
    return 2 * atan(im / re);
}
/* topological charge of a configuration - heisenbergLib.c:712-716 */
double orc_topological_q(const orc_sys *s, const double *sp) {
    double q = 0;
    for (int t = 0; t < s->nTri; t++) {
        int a = s->tri[3 * t], b = s->tri[3 * t + 1], c = s->tri[3 * t + 2];
        q += orc_signed_area(sp + 3 * a, sp + 3 * b, sp + 3 * c, fabs(s->S[a]), fabs(s->S[b]), fabs(s->S[c]));
    }
    return q / REF_PI / 4;
}

/* ------------------------------------------------------------------------------------------ */
/* random helpers of the reference                                                            */
/* ------------------------------------------------------------------------------------------ */
/* generateRandomVec - heisenbergLib.c:95-112 (3 rand per try), xyLib.c:73-88 (2 rand per try) */
static void ref_random_vec(int model, double *n) {
    for (;;) {
        double x = rand() / (double)RAND_MAX - 0.5;
        double y = rand() / (double)RAND_MAX - 0.5;
        double z = 0;
        if (model == 3) z = rand() / (double)RAND_MAX - 0.5;
        double len2 = model == 3 ? (x * x + y * y + z * z) : (x * x + y * y);
        if (len2 > 0.25) continue;
        double len = sqrt(len2);
        n[0] = x / len; n[1] = y / len; n[2] = model == 3 ? z / len : 0;
        return;
    }
}
static int ref_random_site(int N) { /* heisenbergLib.c:443-445 */
    unsigned long long r1 = (unsigned long long)rand();
    unsigned long long r2 = (unsigned long long)rand();
    return (int)((r1 * RAND_MAX + r2) % (unsigned long long)N);
}
/* direction from two Philox words: uniform on S^2 (model 3) or S^1 (model 2) */
static void philox_dir(int model, const uint32_t r[4], int f32, double *n) {
    double u0 = f32 ? u01f(r[0]) : u01(r[0]), u1 = f32 ? u01f(r[1]) : u01(r[1]);
    if (model == 3) {
        double z = 2 * u0 - 1, phi = 2 * M_PI * u1, rr = sqrt(1 - z * z);
        n[0] = rr * cos(phi); n[1] = rr * sin(phi); n[2] = z;
    } else {
        double phi = 2 * M_PI * u0;
        n[0] = cos(phi); n[1] = sin(phi); n[2] = 0;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* initial state - establishLattice heisenbergLib.c:157-172 / xyLib.c:120-132                 */
/* ------------------------------------------------------------------------------------------ */
static void normalize_ref(int model, double *v) { /* heisenbergLib.c:19-25 */
    double len = model == 3 ? sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) : sqrt(v[0] * v[0] + v[1] * v[1]);
    if (len < 1e-5) return;
    v[0] /= len; v[1] /= len;
    if (model == 3) v[2] /= len;
}
void orc_init_spins_ref(const orc_sys *s, double flunc, double *sp) {
    for (int i = 0; i < s->N; i++) {
        double n[3];
        ref_random_vec(s->model, n);
        double *v = sp + 3 * i;
        v[0] = s->S[i]; v[1] = 0; v[2] = 0;
        v[0] += n[0] * flunc; v[1] += n[1] * flunc; v[2] += n[2] * flunc;
        normalize_ref(s->model, v);
        double a = fabs(s->S[i]);
        v[0] *= a; v[1] *= a; v[2] *= a;
        if (s->model == 2) v[2] = 0;
    }
}
/* engine convention for the same formula with a Philox direction (stream INIT) */
void orc_init_spins_philox(const orc_sys *s, double flunc, uint64_t seed, uint32_t replica, int f32, double *sp) {
    for (int i = 0; i < s->N; i++) {
        double n[3] = {0, 0, 0};
        uint32_t r[4];
        rng4(seed, replica, STREAM_INIT, 0, 0, (uint32_t)i, r);
        philox_dir(s->model, r, f32, n);
        double *v = sp + 3 * i;
        v[0] = s->S[i] + n[0] * flunc; v[1] = n[1] * flunc; v[2] = n[2] * flunc;
        normalize_ref(s->model, v);
        double a = fabs(s->S[i]);
        v[0] *= a; v[1] *= a; v[2] *= a;
        if (s->model == 2) v[2] = 0;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* O(n) updates                                                                               */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double energy;
    double tot[3];
    long long attempts, accepted, cluster_sites;
} orc_state;

/* localUpdate - heisenbergLib.c:441-473, xyLib.c:382-409 */
static void on_local_update_ref(const orc_sys *s, double *sp, orc_state *st) {
    int i = ref_random_site(s->N);
    double n[3], tr[3];
    ref_random_vec(s->model, n);
    double s1n = -2 * vdot(s, sp + 3 * i, n);
    tr[0] = n[0] * s1n; tr[1] = n[1] * s1n; tr[2] = s->model == 3 ? n[2] * s1n : 0;
    double corr = orc_delta_energy(s, sp, i, tr);
    st->attempts++;
    if (corr <= 0 || exp(-corr) > rand() / (double)RAND_MAX) {
        st->tot[0] += tr[0]; st->tot[1] += tr[1]; st->tot[2] += tr[2];
        sp[3 * i] += tr[0]; sp[3 * i + 1] += tr[1]; sp[3 * i + 2] += tr[2];
        st->energy += corr;
        st->accepted++;
    }
}

/* one attempt of the colour-sweep variant at site i (same physics, Philox randoms) */
static void on_attempt_philox(const orc_sys *s, double *sp, int i, const uint32_t r[4], int f32, double pAttempt,
                              orc_state *st) {
    if (pAttempt < 1.0 && !((f32 ? u01f(r[3]) : u01(r[3])) < pAttempt)) return;
    double n[3], tr[3];
    philox_dir(s->model, r, f32, n);
    double s1n = -2 * vdot(s, sp + 3 * i, n);
    tr[0] = n[0] * s1n; tr[1] = n[1] * s1n; tr[2] = s->model == 3 ? n[2] * s1n : 0;
    double corr = orc_delta_energy(s, sp, i, tr);
    st->attempts++;
    if (corr <= 0 || exp(-corr) > (f32 ? u01f(r[2]) : u01(r[2]))) {
        sp[3 * i] += tr[0]; sp[3 * i + 1] += tr[1]; sp[3 * i + 2] += tr[2];
        st->accepted++;
    }
}

/* Wolff: expandBlock + blockUpdate - heisenbergLib.c:310-439, xyLib.c:256-380.
 * mode 0: reference (FIFO growth, rand()).  mode 1: Philox variant - identical cluster rule, the
 * uniform of a bond is keyed by its lower-id endpoint and that endpoint's link slot, seed site
 * and plane normal come from stream WSEED; halfMove=0 evaluates the residual with the full move. */
static uint32_t bond_uniform_word(const orc_sys *s, uint64_t seed, uint32_t replica, uint64_t step, int a, int k) {
    /* one Philox word per bond, keyed by (lower site id, higher site id, occurrence, step, replica): engine convention
     * rng_bond().  occurrence = how many earlier link slots of this site lead to the same neighbour (duplicate links of a
     * pair - forceAdd'ed dipole links next to an exchange bond, Lattice.py:298 - are independent bonds, heisenbergLib.c:355-366) */
    int b = s->nbr[(size_t)a * s->maxL + k], occ = 0;
    for (int j = 0; j < k; j++) occ += s->nbr[(size_t)a * s->maxL + j] == b;
    uint32_t lo = (uint32_t)(a < b ? a : b), hi = (uint32_t)(a < b ? b : a);
    uint32_t ctr[4] = {lo, hi, ((uint32_t)STREAM_WBOND << 24) | (((uint32_t)occ >> 2) << 16) | (uint32_t)((step >> 16) & 0xFFFFu),
                       (replica & 0xFFFFu) | ((uint32_t)(step & 0xFFFFu) << 16)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, out[4];
    orc_philox4x32(ctr, key, out);
    return out[occ & 3];
}

static void on_block_update(const orc_sys *s, double *sp, orc_state *st, int mode, uint64_t seed, uint32_t replica,
                            uint64_t step, int f32, int *scratch_i, double *scratch_d) {
    int N = s->N;
    int *inBlock = scratch_i, *isProj = scratch_i + N, *block = scratch_i + 2 * N, *buffer = scratch_i + 3 * N;
    double *trans = scratch_d, *perp = scratch_d + 3 * (size_t)N, *sDotN = scratch_d + 6 * (size_t)N;
    memset(inBlock, 0, sizeof(int) * N);
    memset(isProj, 0, sizeof(int) * N);
    int seedID;
    double n[3];
    uint32_t rs[4] = {0, 0, 0, 0};
    if (mode == 0) {
        seedID = ref_random_site(N);
        ref_random_vec(s->model, n);
    } else {
        rng4(seed, replica, STREAM_WSEED, 0, step, 0, rs);
        seedID = (int)(((uint64_t)rs[3] * (uint64_t)N) >> 32);
        philox_dir(s->model, rs, f32, n);
    }
    block[0] = seedID; buffer[0] = seedID; inBlock[seedID] = 1;
    int begin = 0, end = 0, blockLen = 1;
#define PROJECT(o)                                                                       \
    if (!isProj[o]) {                                                                    \
        sDotN[o] = -vdot(s, sp + 3 * (o), n);                                            \
        for (int c_ = 0; c_ < 3; c_++) {                                                 \
            trans[3 * (o) + c_] = n[c_] * sDotN[o];                                      \
            perp[3 * (o) + c_] = sp[3 * (o) + c_] + trans[3 * (o) + c_];                 \
        }                                                                                \
        isProj[o] = 1;                                                                   \
    }
    while (begin <= end) { /* expandBlock, FIFO */
        int a = buffer[begin++];
        PROJECT(a);
        for (int k = 0; k < s->nlink[a]; k++) {
            int b = s->nbr[(size_t)a * s->maxL + k];
            if (inBlock[b]) continue;
            PROJECT(b);
            double corr = 2 * sDotN[a] * sDotN[b] * ddot(s, n, n, s->J + ((size_t)a * s->maxL + k) * 9);
            if (corr < 0) {
                double u;
                if (mode == 0) u = rand() / (double)RAND_MAX;
                else { uint32_t w = bond_uniform_word(s, seed, replica, step, a, k); u = f32 ? u01f(w) : u01(w); }
                if ((1 - exp(corr)) > u) {
                    block[blockLen++] = b; inBlock[b] = 1; buffer[++end] = b;
                }
            }
        }
    }
    /* residual ("anisotropy") energy of the reflection - heisenbergLib.c:403-418 */
    double res = 0;
    for (int q = 0; q < blockLen; q++) {
        int a = block[q];
        for (int k = 0; k < s->nlink[a]; k++) {
            int b = s->nbr[(size_t)a * s->maxL + k];
            const double *J = s->J + ((size_t)a * s->maxL + k) * 9;
            PROJECT(b); /* the reference reads perpenSpin of possibly un-projected neighbours (zeros from a
                           previous step or stale values); projecting here is the intended quantity and is
                           identical whenever the reference's value is defined by this step */
            double src = sDotN[a] * ddot(s, n, perp + 3 * b, J);
            res += src;
            if (inBlock[b]) res += sDotN[b] * ddot(s, perp + 3 * a, n, J);
            else res += src;
        }
    }
    for (int q = 0; q < blockLen; q++) {
        int a = block[q];
        double tr[3] = {trans[3 * a], trans[3 * a + 1], trans[3 * a + 2]};
        if (!s->wolffHalfMove) { tr[0] *= 2; tr[1] *= 2; tr[2] *= 2; }
        /* getDeltaOnsiteEnergy without the bond part */
        const double *v = sp + 3 * a, *D = s->D + 3 * a;
        double s1x = v[0] + tr[0], s1y = v[1] + tr[1], s1z = v[2] + tr[2];
        if (s->model == 3)
            res += D[0] * (s1x * s1x - v[0] * v[0]) + D[1] * (s1y * s1y - v[1] * v[1]) +
                   D[2] * (s1z * s1z - v[2] * v[2]) - s->h * tr[2];
        else
            res += D[0] * (s1x * s1x - v[0] * v[0]) + D[1] * (s1y * s1y - v[1] * v[1]) - s->h * tr[0];
    }
#undef PROJECT
    st->attempts++;
    double u = mode == 0 ? 0 : (f32 ? u01f(rs[2]) : u01(rs[2]));
    if (res <= 0 || exp(-res) > (mode == 0 ? rand() / (double)RAND_MAX : u)) {
        for (int q = 0; q < blockLen; q++) {
            int a = block[q];
            for (int c = 0; c < 3; c++) {
                double t2 = trans[3 * a + c] * 2;
                sp[3 * a + c] += t2;
                st->tot[c] += t2;
            }
        }
        st->energy = orc_total_energy(s, sp);
        st->accepted++;
        st->cluster_sites += blockLen;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* per-sweep measurement + accumulators - heisenbergLib.c:661-831, xyLib.c:589-758            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double spin_i[3], spin_j[3], spin_ij, totEnergy, E2, M2, M4, M_tmp, MdotM, M_tot;
    double spin_i_r[3], spin_j_r[3], spin_ij_r, totEnergy_r, E2_r;
    double spin_i_z, spin_j_z, spin_tot_z, spin_i_h, spin_j_h, spin_tot_h, topoQ;
} on_acc;

static void majority_spin(const orc_sys *s, const double *sp, int row, double *out) { /* heisenbergLib.c:255-267 */
    double a[3] = {0, 0, 0};
    for (int q = 0; q < s->nC; q++) {
        int o = s->rCl[(size_t)row * s->nC + q];
        a[0] += sp[3 * o]; a[1] += sp[3 * o + 1]; a[2] += sp[3 * o + 2];
    }
    if (s->model == 2) a[2] = 0;
    normalize_ref(s->model, a);
    double S = s->S[s->rOrb[row]]; /* cTimes(&avgSpin,_orb->S): signed S */
    out[0] = a[0] * S; out[1] = a[1] * S; out[2] = a[2] * S;
}

static void on_measure(const orc_sys *s, const double *sp, const orc_state *st, const int *rowOf, on_acc *A,
                       double *gDot, double *g4) {
    int ax = s->model == 3 ? 2 : 0; /* field axis: z (Heis) / x (XY) */
    double dir[3] = {st->tot[0], st->tot[1], s->model == 3 ? st->tot[2] : 0};
    normalize_ref(s->model, dir);
    double si[3] = {0, 0, 0}, sj[3] = {0, 0, 0}, sij = 0, siz = 0, sjz = 0, sih = 0, sjh = 0;
    double nLat = (double)s->nLat;
    for (int j = 0; j < s->nLat; j++) {
        const double *a = sp + 3 * s->pairs[2 * j], *b = sp + 3 * s->pairs[2 * j + 1];
        si[0] += a[0]; si[1] += a[1]; si[2] += a[2];
        sj[0] += b[0]; sj[1] += b[1]; sj[2] += b[2];
        sij += vdot(s, a, b);
        siz += vdot(s, dir, a); sjz += vdot(s, dir, b);
        sih += a[ax]; sjh += b[ax];
    }
    if (s->model == 3) A->topoQ += orc_topological_q(s, sp); /* XY: Q == 0 (xyLib.c:638) */
    A->spin_i_z += siz / nLat; A->spin_j_z += sjz / nLat;
    A->spin_tot_z += vdot(s, dir, st->tot) / nLat;
    A->spin_i_h += sih / nLat; A->spin_j_h += sjh / nLat;
    A->spin_tot_h += st->tot[ax] / nLat;
    double M;
    if (s->model == 3) M = sqrt(vdot(s, st->tot, st->tot)) / nLat;       /* heisenbergLib.c:726 */
    else M = sqrt(vdot(s, si, si)) / nLat;                                /* xyLib.c:654 */
    A->M2 += M * M; A->M4 += M * M * M * M; A->M_tot += M; A->MdotM += A->M_tmp * M; A->M_tmp = M;
    for (int c = 0; c < 3; c++) { A->spin_i[c] += fabs(si[c] / nLat); A->spin_j[c] += fabs(sj[c] / nLat); }
    A->spin_ij += sij / nLat;
    double e_avg = st->energy / s->N;
    A->totEnergy += e_avg; A->E2 += e_avg * e_avg;
    /* block-spin lattice - heisenbergLib.c:748-803 */
    if (s->nR > 0) {
        double ri[3] = {0, 0, 0}, rj[3] = {0, 0, 0}, rij = 0, mi[3], mj[3];
        int ci = 0, cj = 0, cij = 0;
        for (int j = 0; j < s->nLat; j++) {
            int ra = rowOf[s->pairs[2 * j]], rb = rowOf[s->pairs[2 * j + 1]];
            if (ra >= 0) { ci++; majority_spin(s, sp, ra, mi); ri[0] += mi[0]; ri[1] += mi[1]; ri[2] += mi[2]; }
            if (rb >= 0) { cj++; majority_spin(s, sp, rb, mj); rj[0] += mj[0]; rj[1] += mj[1]; rj[2] += mj[2]; }
            if (ra >= 0 && rb >= 0) { cij++; majority_spin(s, sp, ra, mi); majority_spin(s, sp, rb, mj); rij += vdot(s, mi, mj); }
        }
        for (int c = 0; c < 3; c++) { A->spin_i_r[c] += fabs(ri[c] / ci); A->spin_j_r[c] += fabs(rj[c] / cj); }
        A->spin_ij_r += rij / cij;
        double er = 0;
        for (int row = 0; row < s->nR; row++) { /* getCorrEnergy_rnorm: J of the ORIGINAL link slot */
            int o = s->rOrb[row];
            double ms[3], mt[3];
            majority_spin(s, sp, row, ms);
            for (int k = 0; k < s->nlink[o]; k++) {
                int t = s->rNbr[(size_t)row * s->maxL + k];
                if (t < 0 || rowOf[t] < 0) continue; /* reference: UB (odd supercell); skipped here */
                majority_spin(s, sp, rowOf[t], mt);
                er += ddot(s, ms, mt, s->J + ((size_t)o * s->maxL + k) * 9);
            }
        }
        er /= 2;
        for (int row = 0; row < s->nR; row++) er += orc_site_onsite_energy(s, sp, s->rOrb[row]); /* :799 un-renormalised */
        er /= s->nR;
        A->totEnergy_r += er; A->E2_r += er * er;
    }
    /* orbital-group statistics - heisenbergLib.c:806-830 */
    if (gDot) {
        int nG = s->nG;
        double g[(nG + 1) * 3];
        for (int a = 0; a < nG; a++) {
            double t[3] = {0, 0, 0};
            for (int k = 0; k < s->maxG; k++) {
                int o = s->groups[(size_t)a * s->maxG + k];
                if (o < 0) break;
                t[0] += sp[3 * o]; t[1] += sp[3 * o + 1]; t[2] += sp[3 * o + 2];
            }
            g[3 * a] = t[0]; g[3 * a + 1] = t[1]; g[3 * a + 2] = t[2];
        }
        g[3 * nG] = st->tot[0] / nLat; g[3 * nG + 1] = st->tot[1] / nLat; g[3 * nG + 2] = st->tot[2] / nLat;
        for (int a = 0; a <= nG; a++)
            for (int b = 0; b <= nG; b++) {
                double d = vdot(s, g + 3 * a, g + 3 * b);
                gDot[a * (nG + 1) + b] += d;
                if (a == b) g4[a] += d * d;
            }
    }
}

static void recompute_state(const orc_sys *s, const double *sp, orc_state *st) {
    st->energy = orc_total_energy(s, sp);
    st->tot[0] = st->tot[1] = st->tot[2] = 0;
    for (int i = 0; i < s->N; i++) { st->tot[0] += sp[3 * i]; st->tot[1] += sp[3 * i + 1]; st->tot[2] += sp[3 * i + 2]; }
}

/* observables of ONE configuration with the reference's definitions (the engine's measurement
 * kernels are checked against this): out27 as one sweep's contribution (nsweep = 1). */
void orc_observe_on(const orc_sys *s, const double *sp, double *out27, double *groupOut);

/* MCMainFunction - heisenbergLib.c:478-885 / xyLib.c:413-809.
 * update_mode: 0 localUpdate(rand)  1 blockUpdate(rand)  2 Philox colour sweeps  3 Philox Wolff
 * order[N]: colour-major visiting order for mode 2 (ignored otherwise).
 * spins_io: if init_given!=0 the start configuration (else reference init with `flunc`, or the
 *           Philox init in modes 2/3); on return the final configuration.
 * Returns 0; out27 = tuple slots 0..26, frames[spinFrame*N*3], groupOut[(nG+2)(nG+1)]. */
int orc_run(const orc_sys *s, int update_mode, long nthermal, long nsweep, long ninterval, double flunc, int spinFrame,
            const int *order, uint64_t seed, uint32_t replica, int f32, int init_given, double *spins_io,
            double *out27, double *frames, double *groupOut, long long *counters) {
    int N = s->N;
    double *sp = spins_io;
    if (!init_given) {
        if (update_mode <= 1) orc_init_spins_ref(s, flunc, sp);
        else orc_init_spins_philox(s, flunc, seed, replica, f32, sp);
    }
    int *rowOf = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));
    for (int i = 0; i < N; i++) rowOf[i] = -1;
    for (int r = 0; r < s->nR; r++) rowOf[s->rOrb[r]] = r;
    int *scr_i = (int *)malloc(sizeof(int) * 4 * (size_t)(N > 0 ? N : 1));
    double *scr_d = (double *)malloc(sizeof(double) * 7 * (size_t)(N > 0 ? N : 1));
    orc_state st;
    memset(&st, 0, sizeof st);
    recompute_state(s, sp, &st);
    uint64_t sweepCtr = 0;
    /* mapping of (ninterval single-site attempts) onto colour sweeps, engine convention: floor(ninterval/N) whole sweeps, then -
     * for a remainder - one sweep in which every site attempts with probability (ninterval mod N)/N */
    long nfull = 0;
    double pPart = 0.0;
    if (update_mode == 2) {
        nfull = (long)(ninterval / N);
        pPart = (double)(ninterval % N) / (double)N;
    }
#define DO_UPDATES(count)                                                                                   \
    for (long long q_ = 0; q_ < (count); q_++) {                                                            \
        if (update_mode == 0) on_local_update_ref(s, sp, &st);                                              \
        else if (update_mode == 1) on_block_update(s, sp, &st, 0, 0, 0, 0, 0, scr_i, scr_d);                \
        else if (update_mode == 3) { on_block_update(s, sp, &st, 1, seed, replica, sweepCtr, f32, scr_i, scr_d); sweepCtr++; } \
    }
#define DO_SWEEPS(count, pa)                                                                                \
    for (long long q_ = 0; q_ < (count); q_++) {                                                            \
        for (int p_ = 0; p_ < N; p_++) {                                                                    \
            int i_ = order[p_];                                                                             \
            uint32_t r_[4];                                                                                 \
            site_words(s, seed, replica, sweepCtr, (uint32_t)i_, (pa) < 1.0, r_);                           \
            on_attempt_philox(s, sp, i_, r_, f32, (pa), &st);                                               \
        }                                                                                                   \
        sweepCtr++;                                                                                         \
    }
#define DO_INTERVALS(count)                                                                                 \
    for (long long v_ = 0; v_ < (count); v_++) {                                                            \
        DO_SWEEPS(nfull, 1.0);                                                                              \
        if (pPart > 0.0) DO_SWEEPS(1, pPart);                                                               \
    }
    if (update_mode == 2) {
        /* thermalisation: nthermal measurement-intervals worth of updates */
        DO_INTERVALS(nthermal);
        recompute_state(s, sp, &st);
    } else {
        DO_UPDATES((long long)ninterval * (long long)nthermal); /* heisenbergLib.c:614-620 */
    }
    on_acc A;
    memset(&A, 0, sizeof A);
    int nG = s->nG;
    double *gDot = NULL, *g4 = NULL;
    if (nG > 0 || groupOut) {
        gDot = (double *)calloc((size_t)(nG + 1) * (nG + 1), sizeof(double));
        g4 = (double *)calloc((size_t)(nG + 1), sizeof(double));
    }
    long per = nsweep;
    int iFrame = 0;
    if (spinFrame > 0) per = nsweep / spinFrame;
    if (per < 1) per = 1;
    for (long isweep = 0; isweep < nsweep; isweep++) {
        if (update_mode == 2) { DO_INTERVALS(1); recompute_state(s, sp, &st); }
        else { DO_UPDATES(ninterval); if (update_mode == 3) recompute_state(s, sp, &st); }
        if (spinFrame > 0 && isweep % per == 0 && iFrame < spinFrame) { /* :664-675 (+cap) */
            memcpy(frames + (size_t)iFrame * N * 3, sp, sizeof(double) * 3 * (size_t)N);
            iFrame++;
        }
        on_measure(s, sp, &st, rowOf, &A, gDot, g4);
    }
    double ns = (double)nsweep;
    double U4 = (A.M2 / ns) * (A.M2 / ns) / (A.M4 / ns);
    /* heisenbergLib.c:834 writes M_tot/nsweep*M_tot/nsweep, xyLib.c:761 (M_tot/nsweep)*(M_tot/nsweep) */
    double autoCorr = s->model == 3 ? (A.MdotM / ns - A.M_tot / ns * A.M_tot / ns) : (A.MdotM / ns - (A.M_tot / ns) * (A.M_tot / ns));
    double *o = out27;
    o[0] = A.spin_i[0] / ns; o[1] = A.spin_i[1] / ns; o[2] = s->model == 3 ? A.spin_i[2] / ns : 0;
    o[3] = A.spin_j[0] / ns; o[4] = A.spin_j[1] / ns; o[5] = s->model == 3 ? A.spin_j[2] / ns : 0;
    o[6] = A.spin_ij / ns; o[7] = autoCorr; o[8] = A.totEnergy / ns; o[9] = A.E2 / ns; o[10] = U4;
    o[11] = A.spin_i_r[0] / ns; o[12] = A.spin_i_r[1] / ns; o[13] = s->model == 3 ? A.spin_i_r[2] / ns : 0;
    o[14] = A.spin_j_r[0] / ns; o[15] = A.spin_j_r[1] / ns; o[16] = s->model == 3 ? A.spin_j_r[2] / ns : 0;
    o[17] = A.spin_ij_r / ns; o[18] = A.totEnergy_r / ns; o[19] = A.E2_r / ns;
    o[20] = A.spin_i_z / ns; o[21] = A.spin_j_z / ns; o[22] = A.spin_tot_z / ns;
    o[23] = A.spin_i_h / ns; o[24] = A.spin_j_h / ns; o[25] = A.spin_tot_h / ns;
    o[26] = A.topoQ / ns;
    if (groupOut && gDot) {
        for (int a = 0; a < (nG + 1) * (nG + 1); a++) groupOut[a] = gDot[a] / ns;
        for (int a = 0; a <= nG; a++) groupOut[(nG + 1) * (nG + 1) + a] = g4[a] / ns;
    }
    if (counters) { counters[0] = st.attempts; counters[1] = st.accepted; counters[2] = st.cluster_sites; }
    free(gDot); free(g4); free(rowOf); free(scr_i); free(scr_d);
    return 0;
}

void orc_observe_on(const orc_sys *s, const double *sp, double *out27, double *groupOut) {
    double *tmp = (double *)malloc(sizeof(double) * 3 * (size_t)s->N);
    memcpy(tmp, sp, sizeof(double) * 3 * (size_t)s->N);
    /* mode 0 with zero updates: nthermal=0, ninterval=0, nsweep=1, start configuration given */
    orc_run(s, 0, 0, 1, 0, 0.0, 0, NULL, 0, 0, 0, 1, tmp, out27, NULL, groupOut, NULL);
    free(tmp);
}

/* ------------------------------------------------------------------------------------------ */
/* Ising - isingLib.c                                                                         */
/* ------------------------------------------------------------------------------------------ */
static double ising_majority(const orc_sys *s, const double *sp, int row, int use_rand, uint64_t seed, uint32_t replica, uint64_t meas) {
    /* getMajoritySpin - isingLib.c:133-150; ties are broken with rand() in the reference */
    double avg = 0;
    for (int q = 0; q < s->nC; q++) avg += sp[s->rCl[(size_t)row * s->nC + q]];
    double a = fabs(sp[s->rOrb[row]]);
    if (avg > 0) return a;
    if (avg < 0) return -a;
    if (use_rand) return (rand() / (double)RAND_MAX > 0.5) ? a : -a;
    uint32_t w[4]; /* engine convention: stream 5 (block-spin ties), counter = (site, measurement index) */
    rng4(seed, replica, 5, 0, meas, (uint32_t)s->rOrb[row], w);
    return u01(w[0]) > 0.5 ? a : -a;
}
static void ising_local_update_ref(const orc_sys *s, double *sp, orc_state *st) { /* isingLib.c:238-254 */
    int i = ref_random_site(s->N);
    double corr = orc_ising_flip_corr(s, sp, i);
    st->attempts++;
    if (corr >= 0) {
        sp[i] *= -1; st->tot[0] += sp[i] * 2; st->energy -= corr; st->accepted++;
    } else if (exp(corr) > rand() / (double)RAND_MAX) {
        sp[i] *= -1; st->tot[0] += sp[i] * 2; st->energy -= corr; st->accepted++;
    }
}
static void ising_attempt_philox(const orc_sys *s, double *sp, int i, uint32_t wAcc, uint32_t wAtt, int f32, double pAtt, orc_state *st) {
    if (pAtt < 1.0 && !((f32 ? u01f(wAtt) : u01(wAtt)) < pAtt)) return;
    double corr = orc_ising_flip_corr(s, sp, i);
    st->attempts++;
    if (corr >= 0 || exp(corr) > (f32 ? u01f(wAcc) : u01(wAcc))) { sp[i] *= -1; st->accepted++; }
}
static void ising_block_update(const orc_sys *s, double *sp, orc_state *st, int mode, uint64_t seed, uint32_t replica,
                               uint64_t step, int f32, int *scr) { /* isingLib.c:165-236 */
    int N = s->N;
    int *inBlock = scr, *block = scr + N, *buffer = scr + 2 * N;
    memset(inBlock, 0, sizeof(int) * N);
    int seedID;
    uint32_t rs[4] = {0, 0, 0, 0};
    if (mode == 0) seedID = ref_random_site(N);
    else { rng4(seed, replica, STREAM_WSEED, 0, step, 0, rs); seedID = (int)(((uint64_t)rs[3] * (uint64_t)N) >> 32); }
    block[0] = seedID; buffer[0] = seedID; inBlock[seedID] = 1;
    int begin = 0, end = 0, blockLen = 1;
    while (begin <= end) {
        int a = buffer[begin++];
        const double *J = isingJ(s, a);
        for (int k = 0; k < s->nlink[a]; k++) {
            int b = s->nbr[(size_t)a * s->maxL + k];
            if (inBlock[b]) continue;
            double corr = J[k] * sp[a] * sp[b];
            if (corr < 0) {
                double u;
                if (mode == 0) u = rand() / (double)RAND_MAX;
                else { uint32_t w = bond_uniform_word(s, seed, replica, step, a, k); u = f32 ? u01f(w) : u01(w); }
                if ((1 - exp(2 * corr)) > u) { block[blockLen++] = b; inBlock[b] = 1; buffer[++end] = b; }
            }
        }
    }
    double dE = 0;
    for (int q = 0; q < blockLen; q++) dE += 2 * s->h * sp[block[q]]; /* getDeltaOnsiteEnergy :129-131 */
    st->attempts++;
    double u = mode == 0 ? 0 : (f32 ? u01f(rs[2]) : u01(rs[2]));
    if (dE <= 0 || exp(-dE) > (mode == 0 ? rand() / (double)RAND_MAX : u)) {
        for (int q = 0; q < blockLen; q++) {
            int a = block[q];
            sp[a] *= -1;
            if (mode != 0 || a < N - 1) st->tot[0] += sp[a] * 2; /* :228 skips the last site (quirk) */
        }
        st->energy = orc_total_energy(s, sp);
        st->accepted++;
        st->cluster_sites += blockLen;
    }
}

/* MCMainFunction - isingLib.c:259-451.  out10 = tuple slots 0..9. */
int orc_run_ising(const orc_sys *s, int update_mode, long nthermal, long nsweep, long ninterval, int spinFrame,
                  const int *order, uint64_t seed, uint32_t replica, int f32, double *spins_io, double *out10,
                  double *frames, long long *counters) {
    int N = s->N;
    double *sp = spins_io; /* initSpin carries the configuration (isingLib.c:27) */
    int *rowOf = (int *)malloc(sizeof(int) * (N > 0 ? N : 1));
    for (int i = 0; i < N; i++) rowOf[i] = -1;
    for (int r = 0; r < s->nR; r++) rowOf[s->rOrb[r]] = r;
    int *scr = (int *)malloc(sizeof(int) * 3 * (size_t)(N > 0 ? N : 1));
    orc_state st;
    memset(&st, 0, sizeof st);
    uint64_t sweepCtr = 0;
    long nfull = 0;
    double pPart = 0.0;
    if (update_mode == 2) { nfull = (long)(ninterval / N); pPart = (double)(ninterval % N) / (double)N; }   /* as in orc_run */
    if (update_mode <= 1) {
        /* isingLib.c:348-350: energy starts at 0 (relative!), totSpin sums `ninterval` entries */
        long lim = ninterval < N ? ninterval : N;
        for (long i = 0; i < lim; i++) st.tot[0] += sp[i];
    } else {
        for (int i = 0; i < N; i++) st.tot[0] += sp[i];
        st.energy = orc_total_energy(s, sp);
    }
#define I_UPDATES(count)                                                                                     \
    for (long long q_ = 0; q_ < (count); q_++) {                                                             \
        if (update_mode == 0) ising_local_update_ref(s, sp, &st);                                            \
        else if (update_mode == 1) ising_block_update(s, sp, &st, 0, 0, 0, 0, 0, scr);                       \
        else if (update_mode == 3) { ising_block_update(s, sp, &st, 1, seed, replica, sweepCtr, f32, scr); sweepCtr++; } \
    }
#define I_SWEEPS(count, pa)                                                                                  \
    for (long long q_ = 0; q_ < (count); q_++) {                                                             \
        for (int p_ = 0; p_ < N; p_++) {                                                                     \
            int i_ = order[p_];                                                                              \
            uint32_t r_[4];                                                                                  \
            site_words(s, seed, replica, sweepCtr, (uint32_t)i_, (pa) < 1.0, r_);                            \
            ising_attempt_philox(s, sp, i_, r_[2], r_[3], f32, (pa), &st);                                   \
        }                                                                                                    \
        sweepCtr++;                                                                                          \
    }
#define I_INTERVALS(count)                                                                                   \
    for (long long v_ = 0; v_ < (count); v_++) {                                                             \
        I_SWEEPS(nfull, 1.0);                                                                                \
        if (pPart > 0.0) I_SWEEPS(1, pPart);                                                                 \
    }
    if (update_mode == 2) { I_INTERVALS(nthermal); }
    else { I_UPDATES((long long)((int)nthermal * (int)ninterval)); } /* :352 int product */
    double spin_i = 0, spin_j = 0, spin_ij = 0, totE = 0, totEr = 0, E2 = 0, E2r = 0;
    if (update_mode <= 1) st.energy = 0; /* :359 */
    double M2 = 0, M4 = 0, M_tmp = 0, MdotM = 0, M_tot = 0, spin_tot = 0;
    long per = nsweep;
    int iFrame = 0;
    if (spinFrame > 0) per = nsweep / spinFrame;
    if (per < 1) per = 1;
    double nLat = (double)s->nLat;
    for (long isw = 0; isw < nsweep; isw++) {
        if (update_mode == 2) { I_INTERVALS(1); }
        else { I_UPDATES(ninterval); }
        if (update_mode >= 2) {
            st.energy = orc_total_energy(s, sp);
            st.tot[0] = 0;
            for (int i = 0; i < N; i++) st.tot[0] += sp[i];
        }
        if (spinFrame > 0 && isw % per == 0 && iFrame < spinFrame) {
            memcpy(frames + (size_t)iFrame * N, sp, sizeof(double) * (size_t)N);
            iFrame++;
        }
        double er = 0; /* :388-391 */
        for (int row = 0; row < s->nR; row++) {
            int o = s->rOrb[row];
            const double *J = isingJ(s, o);
            double corr = 0;
            double ms = ising_majority(s, sp, row, update_mode <= 1, seed, replica, (uint64_t)isw);
            for (int k = 0; k < s->nlink[o]; k++) {
                int t = s->rNbr[(size_t)row * s->maxL + k];
                if (t < 0 || rowOf[t] < 0) continue; /* reference: UB (odd supercell); skipped here */
                double mt = ising_majority(s, sp, rowOf[t], update_mode <= 1, seed, replica, (uint64_t)isw);
                corr += J[k] * ms * mt;
            }
            er += corr / 2 - s->h * sp[o];
        }
        double si = 0, sj = 0, cav = 0;
        for (int j = 0; j < s->nLat; j++) {
            double a = sp[s->pairs[2 * j]], b = sp[s->pairs[2 * j + 1]];
            si += a; sj += b; cav += a * b;
        }
        spin_tot += st.tot[0];
        double M = si / nLat;
        M2 += M * M; M4 += M * M * M * M; M_tot += M; MdotM += M_tmp * M; M_tmp = M;
        spin_i += fabs(si) / nLat; spin_j += fabs(sj) / nLat; spin_ij += cav / nLat;
        double e_avg = st.energy / N, e_r = s->nR > 0 ? er / s->nR : 0;
        totE += e_avg; totEr += e_r; E2 += e_avg * e_avg; E2r += e_r * e_r;
    }
    double ns = (double)nsweep;
    out10[0] = spin_i / ns; out10[1] = spin_j / ns; out10[2] = spin_ij / ns;
    out10[3] = MdotM / ns - (M_tot / ns) * (M_tot / ns);
    out10[4] = totE / ns; out10[5] = E2 / ns; out10[6] = totEr / ns; out10[7] = E2r / ns;
    out10[8] = (M2 / ns) * (M2 / ns) / (M4 / ns);
    out10[9] = spin_tot / ns / nLat;
    if (counters) { counters[0] = st.attempts; counters[1] = st.accepted; counters[2] = st.cluster_sites; }
    free(rowOf); free(scr);
    return 0;
}
