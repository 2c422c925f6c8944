// libmcsolver_b200: host side of the engine and the C ABI of include/mcsolver_b200.h.
//
// Generic (table-driven) path: the legacy MCMainFunction payload (per-site link tables) is
// coloured greedily from its own bond list, reordered colour-major, its exchange tensors and
// (|S|, D) site classes deduplicated into small tables, and everything is made resident in HBM.
// The structured path (structured.cu) takes a compact lattice descriptor instead.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>

#include "kernels_extra.cuh"
#include "kernels_wolff.cuh"
#include "kernels_resident.cuh"
#include "structured.hpp"

namespace mcg {

static thread_local std::string g_last_error;
void set_last_error(const std::string &m) { g_last_error = m; }

template <typename F> static int guarded(F &&f) {
    try {
        f();
        return MCG_OK;
    } catch (const Error &e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc &) {
        set_last_error("host allocation failed");
        return MCG_ERR_ALLOC;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return MCG_ERR_ARG;
    }
}

void *pool_alloc(size_t bytes) {
    int dev = 0;
    MCG_CUDA(cudaGetDevice(&dev));
    static thread_local int configured = -1;
    cudaMemPool_t pool;
    MCG_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    if (configured != dev) {
        uint64_t keep = ~0ull;
        MCG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        configured = dev;
    }
    void *p = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync(&p, bytes ? bytes : 16, pool, 0);
    if (e != cudaSuccess) throw Error(MCG_ERR_ALLOC, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    MCG_CUDA(cudaStreamSynchronize(0));
    return p;
}
void pool_free(void *p) {
    if (p) cudaFreeAsync(p, 0);
}

template <typename T> static T *dalloc(size_t n) {
    return static_cast<T *>(pool_alloc((n ? n : 1) * sizeof(T)));
}
template <typename T> static T *dupload(const std::vector<T> &v) {
    T *p = dalloc<T>(v.size());
    if (!v.empty()) MCG_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}
template <typename real> static void *dupload_real(const std::vector<double> &v) {
    std::vector<real> t(v.begin(), v.end());
    return dupload<real>(t);
}

static void select_device(const mcg_config *cfg, mcg_system *s) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw Error(MCG_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    int dev = cfg->device;
    if (dev < 0) MCG_CUDA(cudaGetDevice(&dev));
    MCG_REQUIRE(dev < n, "device ordinal out of range");
    MCG_CUDA(cudaSetDevice(dev));
    s->device = dev;
}

static GenArgs gen_args(const mcg_system *s) {
    GenArgs a;
    a.N = s->N; a.maxL = s->maxL; a.R = s->R; a.nJ = s->nJ; a.ncls = s->ncls; a.dupLinks = s->dupLinks ? 1 : 0;
    a.nbrp = s->d_nbrp; a.jtype = s->d_jtype; a.Jtab = s->d_Jtab; a.cls = s->d_cls; a.clsS = s->d_clsS; a.clsD = s->d_clsD;
    a.site_of = s->d_site_of; a.spin = s->d_spin; a.beta = s->d_beta; a.field = s->d_field; a.cnt = s->d_cnt;
    a.key = make_rng_key(s->seed);
    a.replica0 = s->replica0;
    return a;
}

// dispatch on (components, precision, full-tensor) -> f.template operator()<NC, real, FULLJ>()
template <typename F> static void dispatch(const mcg_system *s, F &&f) {
    bool d = s->prec == 64, fj = s->fullJ;
    switch (s->NC) {
    case 1: d ? f.template operator()<1, double, false>() : f.template operator()<1, float, false>(); break;
    case 2:
        if (d) fj ? f.template operator()<2, double, true>() : f.template operator()<2, double, false>();
        else fj ? f.template operator()<2, float, true>() : f.template operator()<2, float, false>();
        break;
    default:
        if (d) fj ? f.template operator()<3, double, true>() : f.template operator()<3, double, false>();
        else fj ? f.template operator()<3, float, true>() : f.template operator()<3, float, false>();
    }
}

// ---------------------------------------------------------------------------------------------
// system construction from per-site tables
// ---------------------------------------------------------------------------------------------
static void greedy_colouring(int N, int maxL, const int32_t *nlink, const int32_t *nbr, std::vector<int> &colour, int &C) {
    // symmetric adjacency in CSR form (the reference's tables are symmetric: J on the source,
    // J^T on the target, Lattice.py:260-262; we do not rely on it for race freedom)
    std::vector<int> deg(N + 1, 0);
    for (int i = 0; i < N; i++)
        for (int k = 0; k < nlink[i]; k++) {
            int j = nbr[(size_t)i * maxL + k];
            if (j == i) continue;
            deg[i + 1]++; deg[j + 1]++;
        }
    for (int i = 0; i < N; i++) deg[i + 1] += deg[i];
    std::vector<int> adj(deg[N]), fill(deg.begin(), deg.end() - 1);
    for (int i = 0; i < N; i++)
        for (int k = 0; k < nlink[i]; k++) {
            int j = nbr[(size_t)i * maxL + k];
            if (j == i) continue;
            adj[fill[i]++] = j; adj[fill[j]++] = i;
        }
    colour.assign(N, -1);
    C = 0;
    std::vector<int> mark;   // mark[c] == i  <=> colour c used by a neighbour of i
    for (int i = 0; i < N; i++) {
        for (int e = deg[i]; e < deg[i + 1]; e++) {
            int c = colour[adj[e]];
            if (c >= 0) { if ((int)mark.size() <= c) mark.resize(c + 1, -1); mark[c] = i; }
        }
        int c = 0;
        while (c < (int)mark.size() && mark[c] == i) c++;
        colour[i] = c;
        C = std::max(C, c + 1);
    }
    for (int i = 0; i < N; i++)   // validation: "no two same-colour sites linked"
        for (int e = deg[i]; e < deg[i + 1]; e++)
            if (colour[adj[e]] == colour[i]) throw Error(MCG_ERR_STATE, "internal: colouring is not proper");
}

static void validate_tables(const mcg_tables *t) {
    MCG_REQUIRE(t, "tables is NULL");
    MCG_REQUIRE(t->model >= 1 && t->model <= 3, "model must be 1 (Ising), 2 (XY) or 3 (Heisenberg)");
    MCG_REQUIRE(t->N > 0, "N must be positive");
    MCG_REQUIRE(t->maxL >= 0, "maxL must be non-negative");
    MCG_REQUIRE(t->S && t->nlink && (t->maxL == 0 || (t->J && t->nbr)), "S/nlink/J/nbr must not be NULL");
    MCG_REQUIRE(t->nLat > 0 && t->pairs, "at least one correlated pair is required (nLat>0)");
    for (int i = 0; i < t->N; i++) {
        MCG_REQUIRE(t->nlink[i] >= 0 && t->nlink[i] <= t->maxL, "nlink[i] out of range");
        for (int k = 0; k < t->nlink[i]; k++) {
            int j = t->nbr[(size_t)i * t->maxL + k];
            MCG_REQUIRE(j >= 0 && j < t->N, "linkedOrb index out of range");
        }
    }
    for (int j = 0; j < 2 * t->nLat; j++) MCG_REQUIRE(t->pairs[j] >= 0 && t->pairs[j] < t->N, "corrOrbPair index out of range");
    MCG_REQUIRE(t->nTri >= 0 && (t->nTri == 0 || t->tri), "tri is NULL");
    for (int j = 0; j < 3 * t->nTri; j++) MCG_REQUIRE(t->tri[j] >= 0 && t->tri[j] < t->N, "localCircuits index out of range");
}

static void alloc_replica_state(mcg_system *s, const mcg_config *cfg) {
    s->beta_host.assign(s->R, 1.0);
    s->field_host.assign(s->R, 0.0);
    if (cfg->beta) s->beta_host.assign(cfg->beta, cfg->beta + s->R);
    if (cfg->field) s->field_host.assign(cfg->field, cfg->field + s->R);
    s->d_beta = dupload(s->beta_host);
    s->d_field = dupload(s->field_host);
    s->d_sums = dalloc<double>((size_t)s->R * NSUM);
    s->d_acc = dalloc<double>((size_t)s->R * NACC);
    s->nLabel = s->R;
    s->slot_host.resize(s->R);
    for (int r = 0; r < s->R; r++) s->slot_host[r] = r;
    s->d_slot = dupload(s->slot_host);
    s->d_last = dalloc<double>(4 * (size_t)s->R);
    MCG_CUDA(cudaMemset(s->d_last, 0, sizeof(double) * 4 * s->R));
    s->d_cnt = dalloc<unsigned long long>((size_t)s->R * NCNT);
    MCG_CUDA(cudaMemset(s->d_sums, 0, sizeof(double) * s->R * NSUM));
    MCG_CUDA(cudaMemset(s->d_acc, 0, sizeof(double) * s->R * NACC));
    MCG_CUDA(cudaMemset(s->d_cnt, 0, sizeof(unsigned long long) * s->R * NCNT));
}

static mcg_system *create_from_tables(const mcg_tables *t, const mcg_config *cfg) {
    validate_tables(t);
    MCG_REQUIRE(cfg, "config is NULL");
    MCG_REQUIRE(cfg->precision == 32 || cfg->precision == 64, "precision must be 32 or 64");
    MCG_REQUIRE(cfg->nReplica >= 1, "nReplica must be >= 1");
    std::unique_ptr<mcg_system> sys(new mcg_system());
    mcg_system *s = sys.get();
    select_device(cfg, s);
    s->model = t->model; s->NC = t->model; s->prec = cfg->precision; s->R = cfg->nReplica;
    s->N = t->N; s->maxL = std::max(1, (int)t->maxL);
    s->seed = cfg->seed; s->replica0 = (uint32_t)cfg->replica_offset;
    const int N = s->N, maxL = t->maxL, JW = t->model == 1 ? 1 : 9;
    s->fullJ = t->model != 1 && !t->ignoreOffDiag;

    std::vector<int> colour;
    greedy_colouring(N, maxL, t->nlink, t->nbr, colour, s->C);
    // colour-major permutation (stable: reference id order inside a class)
    s->colourStart.assign(s->C + 1, 0);
    for (int i = 0; i < N; i++) s->colourStart[colour[i] + 1]++;
    for (int c = 0; c < s->C; c++) s->colourStart[c + 1] += s->colourStart[c];
    s->site_of.resize(N); s->pos_of.resize(N);
    {
        std::vector<int> fill(s->colourStart.begin(), s->colourStart.end() - 1);
        for (int i = 0; i < N; i++) { int p = fill[colour[i]]++; s->site_of[p] = i; s->pos_of[i] = p; }
    }
    // deduplicate exchange tensors (index 0 = zero tensor for padding) and (|S|, D) site classes
    std::map<std::vector<double>, int> jmap, cmap;
    std::vector<double> Jtab(JW, 0.0), clsS, clsD;
    jmap[std::vector<double>(JW, 0.0)] = 0;
    std::vector<int32_t> nbrp((size_t)s->maxL * N);
    std::vector<uint16_t> jtype((size_t)s->maxL * N, 0), cls(N);
    for (int p = 0; p < N; p++) {
        int i = s->site_of[p];
        for (int k = 0; k < s->maxL; k++) {
            size_t o = (size_t)k * N + p;
            if (k >= t->nlink[i]) { nbrp[o] = p; jtype[o] = 0; continue; }
            nbrp[o] = s->pos_of[t->nbr[(size_t)i * maxL + k]];
            std::vector<double> J(t->J + ((size_t)i * maxL + k) * JW, t->J + ((size_t)i * maxL + k) * JW + JW);
            if (t->model != 1 && t->ignoreOffDiag) for (int c = 3; c < 9; c++) J[c] = 0.0;
            auto it = jmap.find(J);
            if (it == jmap.end()) {
                MCG_REQUIRE(jmap.size() < 65535, "more than 65535 distinct exchange tensors");
                it = jmap.emplace(J, (int)jmap.size()).first;
                Jtab.insert(Jtab.end(), J.begin(), J.end());
            }
            jtype[o] = (uint16_t)it->second;
        }
        std::vector<double> key = {std::fabs(t->S[i]), t->D ? t->D[3 * i] : 0.0, t->D ? t->D[3 * i + 1] : 0.0,
                                   t->D ? t->D[3 * i + 2] : 0.0};
        if (t->model == 1) key[1] = key[2] = key[3] = 0.0;
        auto ic = cmap.find(key);
        if (ic == cmap.end()) {
            MCG_REQUIRE(cmap.size() < 65535, "more than 65535 distinct (S,D) site classes");
            ic = cmap.emplace(key, (int)cmap.size()).first;
            clsS.push_back(key[0]);
            clsD.insert(clsD.end(), key.begin() + 1, key.end());
        }
        cls[p] = (uint16_t)ic->second;
    }
    s->nJ = (int)jmap.size(); s->ncls = (int)cmap.size();
    {   // several link slots between one pair of sites?  (the Wolff bond uniforms then need the occurrence index)
        std::vector<int32_t> seen;
        for (int i = 0; i < N && !s->dupLinks; i++) {
            seen.assign(t->nbr + (size_t)i * maxL, t->nbr + (size_t)i * maxL + t->nlink[i]);
            std::sort(seen.begin(), seen.end());
            s->dupLinks = std::adjacent_find(seen.begin(), seen.end()) != seen.end();
        }
    }
    s->isoNoOnsite = true;
    for (double dv : clsD) if (dv != 0.0) s->isoNoOnsite = false;
    for (size_t j = 0; j + JW <= Jtab.size() && t->model != 1; j += JW) {
        const double *Jv = Jtab.data() + j;
        bool iso = Jv[0] == Jv[1] && (t->model == 2 || Jv[1] == Jv[2]);
        for (int c = 3; c < 9; c++) if (Jv[c] != 0.0 && (t->model == 3 || c == 3 || c == 6)) iso = false;
        if (!iso) s->isoNoOnsite = false;
    }
    s->S_host.assign(t->S, t->S + N);
    std::vector<double> signS(N);
    for (int p = 0; p < N; p++) signS[p] = t->S[s->site_of[p]];
    // measurement tables in storage positions
    s->nLat = t->nLat; s->nTri = t->model == 3 ? t->nTri : 0;
    std::vector<int32_t> pairs(2 * (size_t)t->nLat), tri(3 * (size_t)s->nTri), mi(N, 0), mj(N, 0);
    s->selfPairs = true;
    for (int j = 0; j < t->nLat; j++) {
        pairs[2 * j] = s->pos_of[t->pairs[2 * j]]; pairs[2 * j + 1] = s->pos_of[t->pairs[2 * j + 1]];
        mi[pairs[2 * j]]++; mj[pairs[2 * j + 1]]++;
        if (pairs[2 * j] != pairs[2 * j + 1]) s->selfPairs = false;
    }
    for (int j = 0; j < 3 * s->nTri; j++) tri[j] = s->pos_of[t->tri[j]];
    s->nG = t->nG; s->maxG = t->maxG; s->nR = t->nR; s->nC = t->nC;
    if (s->nG > 0) MCG_REQUIRE(t->groups && t->maxG >= 1, "orbGroupList is NULL");
    if (s->nR > 0) MCG_REQUIRE(t->rOrb && t->rCl && t->rNbr && t->nC >= 1, "block-spin tables are NULL");
    std::vector<int32_t> groupsP((size_t)s->nG * std::max(1, s->maxG), -1);
    for (size_t i = 0; i < groupsP.size() && s->nG > 0; i++) {
        int o = t->groups[i];
        MCG_REQUIRE(o < N, "orbGroupList index out of range");
        groupsP[i] = o < 0 ? -1 : s->pos_of[o];
    }
    std::vector<int32_t> rowOf(N, -1), rPos(s->nR), rClP((size_t)s->nR * std::max(1, s->nC)), rNbrRow((size_t)s->nR * s->maxL, -1), rNl(s->nR);
    std::vector<int32_t> pairRowI(t->nLat, -1), pairRowJ(t->nLat, -1);
    for (int i = 0; i < s->nR; i++) { MCG_REQUIRE(t->rOrb[i] >= 0 && t->rOrb[i] < N, "rOrb index out of range"); rowOf[t->rOrb[i]] = i; }
    for (int i = 0; i < s->nR; i++) {
        int o = t->rOrb[i];
        rPos[i] = s->pos_of[o];
        rNl[i] = t->nlink[o];
        for (int q = 0; q < s->nC; q++) {
            int m = t->rCl[(size_t)i * s->nC + q];
            MCG_REQUIRE(m >= 0 && m < N, "rOrbCluster index out of range");
            rClP[(size_t)i * s->nC + q] = s->pos_of[m];
        }
        for (int k = 0; k < maxL; k++) {
            int nb = t->rNbr[(size_t)i * maxL + k];
            MCG_REQUIRE(nb < N, "linkedOrb_rnorm index out of range");
            rNbrRow[(size_t)i * s->maxL + k] = nb < 0 ? -1 : rowOf[nb];   // reference: UB for -1 / unchosen (odd or tiny supercells); skipped
        }
    }
    for (int j = 0; j < t->nLat && s->nR > 0; j++) {
        pairRowI[j] = rowOf[t->pairs[2 * j]]; pairRowJ[j] = rowOf[t->pairs[2 * j + 1]];
        if (pairRowI[j] >= 0) s->rg_ci += 1;
        if (pairRowJ[j] >= 0) s->rg_cj += 1;
        if (pairRowI[j] >= 0 && pairRowJ[j] >= 0) s->rg_cij += 1;
    }

    s->d_nbrp = dupload(nbrp); s->d_jtype = dupload(jtype); s->d_cls = dupload(cls);
    s->d_site_of = dupload(s->site_of); s->d_pos_of = dupload(s->pos_of);
    s->d_pairs = dupload(pairs); s->d_tri = dupload(tri); s->d_mi = dupload(mi); s->d_mj = dupload(mj);
    s->d_signS = dupload(signS);
    if (s->nG > 0) s->d_groups = dupload(groupsP);
    if (s->nR > 0) {
        s->d_rPos = dupload(rPos); s->d_rCl = dupload(rClP); s->d_rNbrRow = dupload(rNbrRow); s->d_rNl = dupload(rNl);
        s->d_pairRowI = dupload(pairRowI); s->d_pairRowJ = dupload(pairRowJ);
    }
    if (s->prec == 64) { s->d_Jtab = dupload_real<double>(Jtab); s->d_clsS = dupload_real<double>(clsS); s->d_clsD = dupload_real<double>(clsD); }
    else { s->d_Jtab = dupload_real<float>(Jtab); s->d_clsS = dupload_real<float>(clsS); s->d_clsD = dupload_real<float>(clsD); }
    s->d_spin = pool_alloc((size_t)s->R * s->NC * N * s->real_size());
    s->d_scratch = dalloc<double>(3 * (size_t)N);
    alloc_replica_state(s, cfg);
    if (s->nR > 0) {
        s->d_ms = dalloc<double>((size_t)s->R * s->nR * 3);
        s->d_rsums = dalloc<double>((size_t)s->R * NRS);
        MCG_CUDA(cudaMemset(s->d_rsums, 0, sizeof(double) * s->R * NRS));
    }
    if (s->nG > 0) {
        size_t n1 = s->nG + 1;
        s->d_gsum = dalloc<double>((size_t)s->R * n1 * 3);
        s->d_gacc = dalloc<double>((size_t)s->R * (n1 + 1) * n1);
        MCG_CUDA(cudaMemset(s->d_gsum, 0, sizeof(double) * s->R * n1 * 3));
        MCG_CUDA(cudaMemset(s->d_gacc, 0, sizeof(double) * s->R * (n1 + 1) * n1));
    }
    MCG_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    // table uploads and accumulator clears above went through the legacy default stream, which a non-blocking stream is
    // not ordered with: drain it once before the first kernel can touch them
    MCG_CUDA(cudaDeviceSynchronize());
    return sys.release();
}

static dim3 grid_for(int n, int R) { return dim3((unsigned)((n + 255) / 256), (unsigned)R, 1); }

static void init_spins(mcg_system *s, double flunc) {
    s->wolffPrimed = false;
    if (s->structured) { structured_init_spins(s, flunc); return; }
    GenArgs a = gen_args(s);
    dispatch(s, [&]<int NC, typename real, bool FJ>() {
        s->launches++; k_init_generic<NC, real><<<grid_for(s->N, s->R), 256, 0, s->stream>>>(a, s->d_signS, flunc);
    });
    MCG_CUDA(cudaGetLastError());
}

static void set_spins(mcg_system *s, int r, const double *spins) {
    s->wolffPrimed = false;
    MCG_REQUIRE(r >= 0 && r < s->R && spins, "bad replica index or NULL spins");
    if (s->structured) { structured_set_spins(s, r, spins); return; }
    size_t n = (size_t)s->N * (s->NC == 1 ? 1 : 3);
    MCG_CUDA(cudaMemcpyAsync(s->d_scratch, spins, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    GenArgs a = gen_args(s);
    dispatch(s, [&]<int NC, typename real, bool FJ>() {
        s->launches++; k_scatter_frame<NC, real><<<grid_for(s->N, 1), 256, 0, s->stream>>>(a, r, s->d_scratch);
    });
    MCG_CUDA(cudaGetLastError());
    MCG_CUDA(cudaStreamSynchronize(s->stream));
}

static void get_spins(mcg_system *s, int r, double *spins) {
    MCG_REQUIRE(r >= 0 && r < s->R && spins, "bad replica index or NULL spins");
    if (s->structured) { structured_get_spins(s, r, spins); return; }
    size_t n = (size_t)s->N * (s->NC == 1 ? 1 : 3);
    GenArgs a = gen_args(s);
    dispatch(s, [&]<int NC, typename real, bool FJ>() {
        s->launches++; k_gather_frame<NC, real><<<grid_for(s->N, 1), 256, 0, s->stream>>>(a, r, s->d_scratch);
    });
    MCG_CUDA(cudaGetLastError());
    MCG_CUDA(cudaMemcpyAsync(spins, s->d_scratch, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    MCG_CUDA(cudaStreamSynchronize(s->stream));
}

// raw per-sweep sums of the resident configuration (all replicas) into d_sums
static void launch_measure_sums(mcg_system *s, int siteRep, double *eb, double *eo) {
    if (s->structured) { structured_measure_sums(s); return; }
    GenArgs a = gen_args(s);
    dispatch(s, [&]<int NC, typename real, bool FJ>() {
        s->launches += 2; k_measure_generic<NC, real, FJ><<<grid_for(s->N, s->R), 256, 0, s->stream>>>(a, s->d_mi, s->d_mj, s->d_sums, siteRep, eb, eo);
        k_pairs_generic<NC, real><<<grid_for(s->nLat, s->R), 256, 0, s->stream>>>(s->N, s->nLat, s->d_pairs, s->d_spin, s->d_sums);
        if constexpr (NC == 3)
            if (s->nTri > 0) s->launches++, k_topo_generic<real><<<grid_for(s->nTri, s->R), 256, 0, s->stream>>>(a, s->nTri, s->d_tri, s->d_sums);
    });
    MCG_CUDA(cudaGetLastError());
}

static void launch_extras(mcg_system *s) {
    if (s->structured || (s->nR == 0 && s->nG == 0)) return;
    GenArgs a = gen_args(s);
    if (s->nR > 0) {
        RgArgs g;
        g.nR = s->nR; g.nC = s->nC; g.nLat = s->nLat; g.rPos = s->d_rPos; g.rCl = s->d_rCl; g.rNbrRow = s->d_rNbrRow; g.rNl = s->d_rNl;
        g.pairRowI = s->d_pairRowI; g.pairRowJ = s->d_pairRowJ; g.ms = s->d_ms; g.rsums = s->d_rsums; g.meas = s->measCtr;
        dispatch(s, [&]<int NC, typename real, bool FJ>() {
            k_rg_majority<NC, real><<<grid_for(s->nR, s->R), 256, 0, s->stream>>>(a, g, s->d_signS);
            k_rg_sums<NC, real, FJ><<<grid_for(std::max(s->nR, s->nLat), s->R), 256, 0, s->stream>>>(a, g);
        });
        s->launches += 2;
    }
    if (s->nG > 0 && s->model != MCG_ISING) {
        dispatch(s, [&]<int NC, typename real, bool FJ>() {
            k_group_sums<NC, real><<<dim3(s->nG, s->R), 256, 0, s->stream>>>(a, s->nG, s->maxG, s->d_groups, s->d_gsum);
        });
        s->launches++;
    }
    k_extra_finalize<<<(s->R + 63) / 64, 64, 0, s->stream>>>(s->model, s->R, s->nLat, s->nR, s->rg_ci, s->rg_cj, s->rg_cij, s->nG,
                                                              s->d_sums, s->d_rsums, s->d_gsum, s->d_acc, s->d_gacc, s->d_slot);
    s->launches++;
    s->measCtr++;
    MCG_CUDA(cudaGetLastError());
}

static void measure(mcg_system *s) {
    launch_measure_sums(s, -1, nullptr, nullptr);
    launch_extras(s);
    s->launches++;
    k_finalize_sweep<<<(s->R + 63) / 64, 64, 0, s->stream>>>(s->model, s->R, s->normN(), s->normLat(), s->d_sums, s->d_acc, s->d_slot, s->d_last);
    MCG_CUDA(cudaGetLastError());
}

static void energy(mcg_system *s, int r, double *Etot, double *eb, double *eo) {
    MCG_REQUIRE(r >= 0 && r < s->R, "bad replica index");
    MCG_REQUIRE((eb == nullptr) == (eo == nullptr), "per-site energy arrays must both be given or both be NULL");
    MCG_REQUIRE(!(s->structured && eb), "per-site energies are only available on table-built systems");
    MCG_CUDA(cudaMemsetAsync(s->d_sums, 0, sizeof(double) * s->R * NSUM, s->stream));
    double *deb = eb ? s->d_scratch : nullptr, *deo = eb ? s->d_scratch + s->N : nullptr;
    launch_measure_sums(s, r, deb, deo);
    double E = 0;
    MCG_CUDA(cudaMemcpyAsync(&E, s->d_sums + (size_t)r * NSUM + SUM_E, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (eb) {
        MCG_CUDA(cudaMemcpyAsync(eb, deb, sizeof(double) * s->N, cudaMemcpyDeviceToHost, s->stream));
        MCG_CUDA(cudaMemcpyAsync(eo, deo, sizeof(double) * s->N, cudaMemcpyDeviceToHost, s->stream));
    }
    MCG_CUDA(cudaMemsetAsync(s->d_sums, 0, sizeof(double) * s->R * NSUM, s->stream));
    MCG_CUDA(cudaStreamSynchronize(s->stream));
    if (Etot) *Etot = E;
}

static void metropolis_sweeps(mcg_system *s, int64_t n, double pAtt) {
    MCG_REQUIRE(n >= 0, "negative sweep count");
    MCG_REQUIRE(pAtt > 0.0 && pAtt <= 1.0, "pAttempt must be in (0,1]");
    s->wolffPrimed = false;
    if (s->structured) { structured_sweeps(s, n, pAtt, false); return; }
    GenArgs a = gen_args(s);
    for (int64_t it = 0; it < n; it++) {
        for (int c = 0; c < s->C; c++) {
            int cb = s->colourStart[c], ce = s->colourStart[c + 1];
            if (ce == cb) continue;
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (s->profilePasses) { MCG_CUDA(cudaEventCreate(&e0)); MCG_CUDA(cudaEventCreate(&e1)); MCG_CUDA(cudaEventRecord(e0, s->stream)); }
            s->launches++;
            dispatch(s, [&]<int NC, typename real, bool FJ>() {
                k_metro_generic<NC, real, FJ><<<grid_for(ce - cb, s->R), 256, 0, s->stream>>>(a, cb, ce, s->sweepCtr, (real)pAtt);
            });
            if (s->profilePasses) { MCG_CUDA(cudaEventRecord(e1, s->stream)); s->passEvents.emplace_back(e0, e1); }
        }
        s->sweepCtr++;
    }
    MCG_CUDA(cudaGetLastError());
}

// frontier/global hybrid of the per-phase Wolff path: 0 = off, 1 = adaptive, 2 = frontier always tried first, 3 = hybrid sequence
// with the frontier kernel declining every step (exercises the hybrid bookkeeping of the global passes alone)
static int wolff_hybrid_mode(const mcg_system *s) {
    if (const char *e = getenv("MCG_WOLFF_FRONTIER")) return atoi(e);
    return s->N >= 65536 ? 1 : 0;    // below that a global pass is a handful of microseconds
}

static void wolff_steps(mcg_system *s, int64_t n) {
    MCG_REQUIRE(n >= 0, "negative step count");
    const size_t RN = (size_t)s->R * s->N;
    if (!s->d_parent) {
        s->d_parent = (int32_t *)pool_alloc(2 * RN * sizeof(int32_t));
        s->d_proj = pool_alloc(2 * RN * s->real_size());
        s->d_wres = dalloc<double>(2 * (size_t)s->R);
        s->wolffPrimed = false;
    }
    // residual energy of a reflection vanishes identically for isotropic exchange without D and field
    bool anyField = false;
    for (double h : s->field_host) anyField = anyField || h != 0.0;
    const bool needResidual = anyField || (s->model != MCG_ISING && !s->isoNoOnsite);
    WolffArgs w;
    w.wres = s->d_wres; w.N = s->N; w.R = s->R; w.spin = s->d_spin;
    w.beta = s->d_beta; w.field = s->d_field; w.cnt = s->d_cnt; w.key = make_rng_key(s->seed); w.replica0 = s->replica0;
    const int hyb = wolff_hybrid_mode(s);
    if (hyb) {
        if (!s->d_wmode) {
            const char *e = getenv("MCG_WOLFF_FRONTIER_CAP");
            s->wolffCap = std::max(4, std::min(s->N, e ? atoi(e) : 32768));
            s->d_wmode = dalloc<int32_t>((size_t)s->R * WM_N);
            s->d_wstamp = (uint32_t *)pool_alloc(RN * sizeof(uint32_t));
            s->d_wqueue = (int32_t *)pool_alloc((size_t)s->R * s->wolffCap * sizeof(int32_t));
            s->d_hparent = (int32_t *)pool_alloc(2 * RN * sizeof(int32_t));
            MCG_CUDA(cudaMemsetAsync(s->d_wmode, 0, sizeof(int32_t) * s->R * WM_N, s->stream));
            MCG_CUDA(cudaMemsetAsync(s->d_wstamp, 0, RN * sizeof(uint32_t), s->stream));
            // half 0 of every replica's forest pair starts as the identity (positions are replica-local)
            for (int r = 0; r < s->R; r++) k_wolff_identity<<<148 * 4, 256, 0, s->stream>>>(s->d_hparent + (size_t)r * s->N, (size_t)s->N);
            s->wolffTag = 0;
            s->wolffPrimed = false;
        }
        w.mode = s->d_wmode; w.stamp = s->d_wstamp; w.queue = s->d_wqueue; w.cap = s->wolffCap;
        w.parent = s->d_hparent; w.parentNext = nullptr; w.proj = s->d_proj; w.projNext = nullptr;
        const int force = hyb == 2 ? 1 : hyb == 3 ? 2 : 0;
        for (int64_t it = 0; it < n; it++) {
            w.step = s->wolffCtr++;
            if (++s->wolffTag == 0xFFFFFFFFu) {   // stamps of 2^32 steps ago must not read as this step's
                MCG_CUDA(cudaMemsetAsync(s->d_wstamp, 0, RN * sizeof(uint32_t), s->stream));
                s->wolffTag = 1;
                s->wolffPrimed = false;
            }
            w.tag = s->wolffTag;
            w.hostPrimed = s->wolffPrimed ? 1 : 0;
            if (s->structured) s->launches += structured_wolff_step(s, w, s->wolffPrimed, needResidual, force);
            else {
                GenArgs a = gen_args(s);
                dispatch(s, [&]<int NC, typename real, bool FJ>() {
                    TableTopo<NC, real> topo{a, s->d_pos_of};
                    s->launches += wolff_launch_hybrid<NC, real, FJ>(topo, w, s->stream, s->maxL, needResidual, force, 148 * 8);
                });
            }
            s->wolffPrimed = true;
        }
        MCG_CUDA(cudaGetLastError());
        return;
    }
    for (int64_t it = 0; it < n; it++) {
        w.step = s->wolffCtr++;
        const int b = (int)(w.step & 1);
        w.parent = s->d_parent + (size_t)b * RN; w.parentNext = s->d_parent + (size_t)(1 - b) * RN;
        w.proj = (char *)s->d_proj + (size_t)b * RN * s->real_size(); w.projNext = (char *)s->d_proj + (size_t)(1 - b) * RN * s->real_size();
        if (s->structured) s->launches += structured_wolff_step(s, w, s->wolffPrimed, needResidual, 0);
        else {
            GenArgs a = gen_args(s);
            dispatch(s, [&]<int NC, typename real, bool FJ>() {
                TableTopo<NC, real> topo{a, s->d_pos_of};
                s->launches += wolff_launch_step<NC, real, FJ>(topo, w, s->stream, s->wolffPrimed, needResidual);
            });
        }
        s->wolffPrimed = true;
    }
    MCG_CUDA(cudaGetLastError());
}

static void capture_frame(mcg_system *s, int r, double *dst) {
    // synchronous: frames are rare (spinFrame per run) and the staging buffer is reused
    get_spins(s, r, dst);
}

// Small table-built lattices: the whole loop runs inside k_resident (kernels_resident.cuh), one block per replica.
static int resident_max_sites() {
    if (getenv("MCG_NO_RESIDENT")) return 0;
    const char *e = getenv("MCG_RESIDENT_MAXN");
    return e ? atoi(e) : 4096;   // measured crossover with the launch-per-phase path (scripts/probe_resident.py)
}

// blocks per replica of the cooperative variant (0 = not applicable: too small to split, too many replicas to co-reside)
static int coop_blocks_per_replica(mcg_system *s) {
    if (getenv("MCG_NO_COOP")) return 0;
    const char *e = getenv("MCG_COOP_MAXN");
    if (s->N > (e ? atoi(e) : 131072)) return 0;
    int dev = 0, sms = 0, coop = 0, perSM = 0;
    MCG_CUDA(cudaGetDevice(&dev));
    MCG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MCG_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return 0;
    dispatch(s, [&]<int NC, typename real, bool FJ>() {
        MCG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_resident_coop<NC, real, FJ>, RES_THREADS, 0));
    });
    const int capacity = sms * perSM;
    const char *ps = getenv("MCG_COOP_SITES");   // sites per block aimed at (tests force many blocks on tiny lattices)
    const int per = std::max(1, ps ? atoi(ps) : RES_THREADS * 4);
    int B = std::min(capacity / std::max(1, s->R), (s->N + per - 1) / per);
    return B >= 2 ? B : 0;
}

static void run_resident(mcg_system *s, int algorithm, int64_t thermalUpdates, int64_t perSweep, double pAtt, int64_t nsweep, int spinFrame,
                         double *frames, int coopB) {
    if (!s->d_colourStart) {
        s->d_colourStart = dalloc<int>(s->colourStart.size());
        MCG_CUDA(cudaMemcpy(s->d_colourStart, s->colourStart.data(), s->colourStart.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    bool needResidual = false;
    if (algorithm == MCG_WOLFF) {
        bool anyField = false;
        for (double h : s->field_host) anyField = anyField || h != 0.0;
        needResidual = anyField || (s->model != MCG_ISING && !s->isoNoOnsite);
    }
    if (coopB == 0 || algorithm != MCG_WOLFF) s->wolffPrimed = false;   // single-block kernel: forest in shared memory, global buffers not prepared
    if (coopB > 0 && algorithm == MCG_WOLFF && !s->d_parent) {
        const size_t RN = (size_t)s->R * s->N;
        s->d_parent = (int32_t *)pool_alloc(2 * RN * sizeof(int32_t));
        s->d_proj = pool_alloc(2 * RN * s->real_size());
        s->d_wres = dalloc<double>(2 * (size_t)s->R);
        s->wolffPrimed = false;
    }
    const size_t fsz = (size_t)s->N * (s->NC == 1 ? 1 : 3);
    double *d_frames = nullptr;
    if (spinFrame > 0) {
        d_frames = dalloc<double>((size_t)s->R * spinFrame * fsz);
        MCG_CUDA(cudaMemsetAsync(d_frames, 0, sizeof(double) * s->R * spinFrame * fsz, s->stream));   // frames never reached stay zero
    }

    GenArgs a = gen_args(s);
    RgArgs g;
    g.nR = s->nR; g.nC = s->nC; g.nLat = s->nLat; g.rPos = s->d_rPos; g.rCl = s->d_rCl; g.rNbrRow = s->d_rNbrRow; g.rNl = s->d_rNl;
    g.pairRowI = s->d_pairRowI; g.pairRowJ = s->d_pairRowJ; g.ms = s->d_ms; g.rsums = s->d_rsums; g.meas = 0;
    WolffArgs w;
    w.parent = w.parentNext = nullptr; w.proj = w.projNext = nullptr;
    w.wres = s->d_wres; w.N = s->N; w.R = s->R; w.spin = s->d_spin; w.step = 0;
    w.beta = s->d_beta; w.field = s->d_field; w.cnt = s->d_cnt; w.key = make_rng_key(s->seed); w.replica0 = s->replica0;
    ResidentPlan P;
    P.algorithm = algorithm; P.model = s->model; P.pAtt = pAtt; P.C = s->C; P.colourStart = s->d_colourStart;
    P.needResidual = needResidual ? 1 : 0;
    P.spinFrame = spinFrame; P.per = spinFrame > 0 ? std::max<int64_t>(1, nsweep / spinFrame) : 1; P.frames = d_frames;
    P.mi = s->d_mi; P.mj = s->d_mj; P.pairs = s->d_pairs; P.tri = s->d_tri; P.nLat = s->nLat; P.nTri = s->nTri;
    P.nG = s->nG; P.maxG = s->maxG; P.groups = s->d_groups; P.gsum = s->d_gsum; P.gacc = s->d_gacc;
    P.rg_ci = s->rg_ci; P.rg_cj = s->rg_cj; P.rg_cij = s->rg_cij; P.signS = s->d_signS;
    P.acc = s->d_acc; P.last = s->d_last; P.slot = s->d_slot;
    const bool extras = s->nR > 0 || s->nG > 0;
    // threads per block: about one site of the largest colour class per thread (barriers and reductions scale with warps)
    int biggest = 1;
    for (int c = 0; c < s->C; c++) biggest = std::max(biggest, s->colourStart[c + 1] - s->colourStart[c]);
    if (algorithm == MCG_WOLFF) biggest = s->N;
    int nthreads = 64;
    while (nthreads < biggest && nthreads < RES_THREADS) nthreads <<= 1;
    CoopPlan Q;
    Q.B = coopB; Q.parentBase = s->d_parent; Q.projBase = (char *)s->d_proj; Q.sums = s->d_sums; Q.rsums = s->d_rsums; Q.primed = 0;
    w.wres = s->d_wres;

    auto launch = [&](int64_t thermal, int64_t i0, int64_t n) {
        P.thermal = thermal; P.perSweep = perSweep; P.nsweep = n; P.i0 = i0;
        P.sweep0 = s->sweepCtr; P.step0 = s->wolffCtr; P.meas0 = s->measCtr;
        if (coopB > 0) {
            Q.primed = s->wolffPrimed ? 1 : 0;
            const int32_t *posOf = s->d_pos_of;
            void *args[] = {(void *)&a, (void *)&g, (void *)&w, (void *)&posOf, (void *)&P, (void *)&Q};
            dispatch(s, [&]<int NC, typename real, bool FJ>() {
                MCG_CUDA(cudaLaunchCooperativeKernel((const void *)k_resident_coop<NC, real, FJ>, dim3((unsigned)(s->R * coopB)), dim3(RES_THREADS), args, 0, s->stream));
            });
        } else {
            const size_t shm = algorithm == MCG_WOLFF ? resident_wolff_smem(s->N, s->real_size()) : 0;
            dispatch(s, [&]<int NC, typename real, bool FJ>() {
                if (shm > 40 * 1024) MCG_CUDA(cudaFuncSetAttribute(k_resident<NC, real, FJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
                k_resident<NC, real, FJ><<<s->R, nthreads, shm, s->stream>>>(a, g, w, s->d_pos_of, P);
            });
        }
        MCG_CUDA(cudaGetLastError());
        s->launches++;
        const int64_t updates = thermal + n * perSweep;
        if (algorithm == MCG_METROPOLIS) s->sweepCtr += updates;
        else { s->wolffCtr += updates; if (coopB > 0 && updates > 0) s->wolffPrimed = true; }
        if (extras) s->measCtr += n;
    };
    const int64_t chunk = 1 << 16;   // bounds the run time of one launch
    for (int64_t done = 0; done < thermalUpdates; done += chunk) launch(std::min(chunk, thermalUpdates - done), 0, 0);
    const int64_t mchunk = std::max<int64_t>(1, chunk / std::max<int64_t>(1, perSweep));
    for (int64_t done = 0; done < nsweep; done += mchunk) launch(0, done, std::min(mchunk, nsweep - done));
    if (spinFrame > 0) {
        MCG_CUDA(cudaMemcpyAsync(frames, d_frames, sizeof(double) * s->R * spinFrame * fsz, cudaMemcpyDeviceToHost, s->stream));
    }
    MCG_CUDA(cudaStreamSynchronize(s->stream));
    pool_free(d_frames);
}

static void run(mcg_system *s, int algorithm, int64_t nthermal, int64_t nsweep, int64_t ninterval, int spinFrame, double *frames) {
    MCG_REQUIRE(algorithm == MCG_METROPOLIS || algorithm == MCG_WOLFF, "algorithm must be 0 (Metropolis) or 1 (Wolff)");
    MCG_REQUIRE(nthermal >= 0 && nsweep >= 1 && ninterval >= 0, "need nthermal>=0, nsweep>=1, ninterval>=0");
    MCG_REQUIRE(spinFrame >= 0 && (spinFrame == 0 || frames), "frames buffer is NULL");
    // Metropolis: `ninterval` single-site attempts between two measurements (heisenbergLib.c:614-620) = floor(ninterval / N) whole
    // colour sweeps plus, for the remainder, one sweep in which every site attempts with probability (ninterval mod N) / N.
    // nsub sweeps per interval, the last of them with attempt probability pAtt.
    int64_t nsub = 1, nfull = 0;
    double pAtt = 1.0;
    if (algorithm == MCG_METROPOLIS) {
        const int64_t Nw = s->normN();     // sites of the whole lattice (a slab holds a part of it plus ghosts)
        nfull = ninterval / Nw;
        const int64_t rem = ninterval % Nw;
        if (rem > 0) pAtt = (double)rem / (double)Nw;
        nsub = nfull + (rem > 0 ? 1 : 0);
    }
    const bool partialLast = pAtt < 1.0;
    if (!s->structured && !s->profilePasses) {
        const int64_t per = algorithm == MCG_METROPOLIS ? nsub : ninterval;
        if (s->N <= resident_max_sites()) { run_resident(s, algorithm, nthermal * per, per, pAtt, nsweep, spinFrame, frames, 0); return; }
        if (const int B = coop_blocks_per_replica(s)) { run_resident(s, algorithm, nthermal * per, per, pAtt, nsweep, spinFrame, frames, B); return; }
    }
    auto updates = [&](int64_t intervals) {
        if (algorithm != MCG_METROPOLIS) { wolff_steps(s, intervals * ninterval); return; }
        if (nsub == 0) return;
        if (!partialLast) { metropolis_sweeps(s, intervals * nfull, 1.0); return; }
        for (int64_t i = 0; i < intervals; i++) {
            if (nfull > 0) metropolis_sweeps(s, nfull, 1.0);
            metropolis_sweeps(s, 1, pAtt);
        }
    };
    // structured Metropolis: the measurement sums are produced by the last sweep of the interval itself
    const bool fused = s->structured && algorithm == MCG_METROPOLIS && nsub > 0;
    updates(nthermal);
    int64_t per = nsweep, iFrame = 0;
    if (spinFrame > 0) per = std::max<int64_t>(1, nsweep / spinFrame);
    size_t fsz = (size_t)s->N * (s->NC == 1 ? 1 : 3);
    for (int64_t i = 0; i < nsweep; i++) {
        if (fused) {
            s->wolffPrimed = false;
            if (!partialLast) structured_sweeps(s, nfull, 1.0, true);
            else {
                if (nfull > 0) structured_sweeps(s, nfull, 1.0, false);
                structured_sweeps(s, 1, pAtt, true);
            }
        }
        else updates(1);
        if (spinFrame > 0 && i % per == 0 && iFrame < spinFrame) {   // heisenbergLib.c:664-675, capped
            for (int r = 0; r < s->R; r++) capture_frame(s, r, frames + ((size_t)r * spinFrame + iFrame) * fsz);
            iFrame++;
        }
        if (fused) {
            s->launches++;
            k_finalize_sweep<<<(s->R + 63) / 64, 64, 0, s->stream>>>(s->model, s->R, s->normN(), s->normLat(), s->d_sums, s->d_acc, s->d_slot, s->d_last);
            MCG_CUDA(cudaGetLastError());
        } else measure(s);
        if ((i & 255) == 255) MCG_CUDA(cudaStreamSynchronize(s->stream));   // bound the launch queue
    }
    MCG_CUDA(cudaStreamSynchronize(s->stream));
}

// accBase / gaccBase: the accumulator rows to read (the system's own, or the cross-rank sums of a tempering run)
void results_from(mcg_system *s, const double *accBase, const double *gaccBase, int r, double *out, double *groupOut) {
    MCG_REQUIRE(r >= 0 && r < s->nLabel && out, "bad replica/label index or NULL out");
    double A[NACC];
    MCG_CUDA(cudaStreamSynchronize(s->stream));
    MCG_CUDA(cudaMemcpy(A, accBase + (size_t)r * NACC, sizeof(A), cudaMemcpyDeviceToHost));
    double ns = A[ACC_NMEAS];
    if (!(ns > 0)) throw Error(MCG_ERR_STATE, "no measurement has been accumulated yet");
    double U4 = (A[ACC_M2] / ns) * (A[ACC_M2] / ns) / (A[ACC_M4] / ns);                     // heisenbergLib.c:833
    double autoCorr = A[ACC_MDOTM] / ns - (A[ACC_MTOT] / ns) * (A[ACC_MTOT] / ns);          // :834
    if (s->model == MCG_ISING) {   // isingLib.c:435-446
        out[0] = A[ACC_SI] / ns; out[1] = A[ACC_SJ] / ns; out[2] = A[ACC_SIJ] / ns; out[3] = autoCorr;
        out[4] = A[ACC_E] / ns; out[5] = A[ACC_E2] / ns; out[6] = A[ACC_ER] / ns; out[7] = A[ACC_E2R] / ns;
        out[8] = U4; out[9] = A[ACC_STOT] / ns / s->normLat();
        return;
    }
    for (int c = 0; c < 3; c++) { out[c] = A[ACC_SI + c] / ns; out[3 + c] = A[ACC_SJ + c] / ns; }
    out[6] = A[ACC_SIJ] / ns; out[7] = autoCorr; out[8] = A[ACC_E] / ns; out[9] = A[ACC_E2] / ns; out[10] = U4;
    for (int c = 0; c < 3; c++) { out[11 + c] = A[ACC_SIR + c] / ns; out[14 + c] = A[ACC_SJR + c] / ns; }
    out[17] = A[ACC_SIJR] / ns; out[18] = A[ACC_ER] / ns; out[19] = A[ACC_E2R] / ns;
    out[20] = A[ACC_SIZ] / ns; out[21] = A[ACC_SJZ] / ns; out[22] = A[ACC_STZ] / ns;
    out[23] = A[ACC_SIH] / ns; out[24] = A[ACC_SJH] / ns; out[25] = A[ACC_STH] / ns;
    out[26] = A[ACC_Q] / ns;
    if (groupOut) {
        int n = (s->nG + 2) * (s->nG + 1);
        if (gaccBase && r < s->nLabel) {
            MCG_CUDA(cudaMemcpy(groupOut, gaccBase + (size_t)r * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
            for (int i = 0; i < n; i++) groupOut[i] /= ns;
        } else
            for (int i = 0; i < n; i++) groupOut[i] = 0.0;
    }
}

static void results(mcg_system *s, int r, double *out, double *groupOut) { results_from(s, s->d_acc, s->d_gacc, r, out, groupOut); }

// one Metropolis sweep followed by the per-sweep measurement (fused into the colour passes on structured systems)
void measured_sweep(mcg_system *s, double pAtt) {
    if (s->structured) {
        s->wolffPrimed = false;
        structured_sweeps(s, 1, pAtt, true);
        s->launches++;
        k_finalize_sweep<<<(s->R + 63) / 64, 64, 0, s->stream>>>(s->model, s->R, s->normN(), s->normLat(), s->d_sums, s->d_acc, s->d_slot, s->d_last);
    } else {
        metropolis_sweeps(s, 1, pAtt);
        measure(s);
    }
}

void reset_measurements_async(mcg_system *sys) {
    MCG_CUDA(cudaMemsetAsync(sys->d_acc, 0, sizeof(double) * sys->nLabel * NACC, sys->stream));
    MCG_CUDA(cudaMemsetAsync(sys->d_sums, 0, sizeof(double) * sys->R * NSUM, sys->stream));
    MCG_CUDA(cudaMemsetAsync(sys->d_cnt, 0, sizeof(unsigned long long) * sys->R * NCNT, sys->stream));
    if (sys->d_gacc) MCG_CUDA(cudaMemsetAsync(sys->d_gacc, 0, sizeof(double) * sys->nLabel * (sys->nG + 2) * (sys->nG + 1), sys->stream));
    sys->measCtr = 0;
}

}  // namespace mcg

mcg_system::~mcg_system() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);   // nothing of this system may still be running when its pages go back to the pool
    mcg::pool_free(d_spin);
    mcg::pool_free(d_parent);
    mcg::pool_free(d_wstamp);
    mcg::pool_free(d_wqueue);
    mcg::pool_free(d_hparent);
    mcg::pool_free(d_proj);
    d_spin = nullptr; d_parent = nullptr; d_proj = nullptr;
    void *bufs[] = {d_nbrp, d_site_of, d_pos_of, d_pairs, d_tri, d_mi, d_mj, d_jtype, d_cls, d_Jtab, d_clsS, d_clsD, d_spin,
                    d_signS, d_beta, d_field, d_sums, d_acc, d_cnt, d_scratch, d_parent, d_proj, d_wres, d_slot, d_last, d_rPos, d_rCl, d_rNbrRow, d_rNl, d_pairRowI, d_pairRowJ, d_groups, d_ms, d_rsums,
                    d_gsum, d_gacc, d_colourStart, d_wmode};
    for (void *b : bufs) mcg::pool_free(b);
    if (st) mcg::structured_destroy(st);
    if (pt) mcg::pt_destroy(pt);
    if (stream) cudaStreamDestroy(stream);
}

// =============================================================================================
// C ABI
// =============================================================================================
using namespace mcg;

extern "C" {

MCG_API const char *mcg_last_error(void) { return g_last_error.c_str(); }
MCG_API int mcg_version(void) { return 100; }

MCG_API int mcg_device_count(int *count) {
    return guarded([&] {
        MCG_REQUIRE(count, "count is NULL");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) throw Error(MCG_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
        *count = n;
    });
}

MCG_API int mcg_create_tables(const mcg_tables *t, const mcg_config *cfg, mcg_system **out) {
    return guarded([&] {
        MCG_REQUIRE(out, "out is NULL");
        *out = create_from_tables(t, cfg);
    });
}

MCG_API int mcg_create_lattice(const mcg_lattice_desc *d, const mcg_config *cfg, mcg_system **out) {
    return guarded([&] {
        MCG_REQUIRE(out, "out is NULL");
        MCG_REQUIRE(d && cfg, "descriptor/config is NULL");
        MCG_REQUIRE(cfg->precision == 32 || cfg->precision == 64 || cfg->precision == 8, "precision must be 64, 32 or (Ising lattices) 8");
        MCG_REQUIRE(cfg->nReplica >= 1, "nReplica must be >= 1");
        std::unique_ptr<mcg_system> sys(new mcg_system());
        select_device(cfg, sys.get());
        sys->prec = cfg->precision; sys->R = cfg->nReplica; sys->seed = cfg->seed; sys->replica0 = (uint32_t)cfg->replica_offset;
        structured_create(sys.get(), d);
        alloc_replica_state(sys.get(), cfg);
        if (sys->nG > 0) {
            size_t n = (size_t)(sys->nG + 2) * (sys->nG + 1);
            sys->d_gacc = dalloc<double>((size_t)sys->R * n);
            MCG_CUDA(cudaMemset(sys->d_gacc, 0, sizeof(double) * sys->R * n));
        }
        MCG_CUDA(cudaStreamCreateWithFlags(&sys->stream, cudaStreamNonBlocking));
        MCG_CUDA(cudaDeviceSynchronize());   // uploads / clears on the legacy stream are complete before s->stream is used
        *out = sys.release();
    });
}

// One lattice cut into `world` slabs along its first axis, one rank per GPU (structured.cu, last section): desc describes the
// WHOLE lattice, the created system holds this rank's planes plus ghost planes.  Sweeps exchange the boundary planes after every
// colour pass and all-reduce the raw measurement sums, so every rank accumulates the whole lattice's observables; the
// trajectory is the undivided lattice's, bit for bit.
MCG_API int mcg_create_lattice_slab(const mcg_lattice_desc *d, const mcg_config *cfg, int rank, int world, const char *comm_id, mcg_system **out) {
    return guarded([&] {
        MCG_REQUIRE(out, "out is NULL");
        MCG_REQUIRE(d && cfg, "descriptor/config is NULL");
        MCG_REQUIRE(cfg->precision == 32 || cfg->precision == 64, "slab decomposition: precision 32 or 64");
        MCG_REQUIRE(cfg->nReplica >= 1, "nReplica must be >= 1");
        std::unique_ptr<mcg_system> sys(new mcg_system());
        select_device(cfg, sys.get());
        sys->prec = cfg->precision; sys->R = cfg->nReplica; sys->seed = cfg->seed; sys->replica0 = (uint32_t)cfg->replica_offset;
        structured_create_slab(sys.get(), d, rank, world, comm_id);
        alloc_replica_state(sys.get(), cfg);
        if (sys->nG > 0) {
            size_t n = (size_t)(sys->nG + 2) * (sys->nG + 1);
            sys->d_gacc = dalloc<double>((size_t)sys->R * n);
            MCG_CUDA(cudaMemset(sys->d_gacc, 0, sizeof(double) * sys->R * n));
        }
        MCG_CUDA(cudaStreamCreateWithFlags(&sys->stream, cudaStreamNonBlocking));
        MCG_CUDA(cudaDeviceSynchronize());
        *out = sys.release();
    });
}

MCG_API int mcg_jit_check(const mcg_lattice_desc *d, int precision, int *ncompiled, char *report, int report_len) {
    return guarded([&] {
        MCG_REQUIRE(d && ncompiled, "NULL argument");
        MCG_REQUIRE(precision == 32 || precision == 64 || precision == 8, "precision must be 64, 32 or 8");
        std::string rep;
        *ncompiled = structured_jit_check(d, precision, rep);
        if (report && report_len > 0) { std::strncpy(report, rep.c_str(), report_len - 1); report[report_len - 1] = 0; }
        if (*ncompiled < 0) throw Error(MCG_ERR_STATE, "NVRTC compilation failed: " + rep);
    });
}

MCG_API int mcg_destroy(mcg_system *sys) {
    return guarded([&] { delete sys; });
}

#define SYS_GUARD(...)                          \
    return guarded([&] {                        \
        MCG_REQUIRE(sys, "system is NULL");     \
        MCG_CUDA(cudaSetDevice(sys->device));   \
        __VA_ARGS__;                            \
    })

MCG_API int mcg_num_colours(const mcg_system *sys, int *ncolours) {
    return guarded([&] { MCG_REQUIRE(sys && ncolours, "NULL argument"); *ncolours = sys->C; });
}

MCG_API int mcg_rng_layout(const mcg_system *sys, int32_t *stride, int32_t *group) {
    return guarded([&] {
        MCG_REQUIRE(sys && stride && group, "NULL argument");
        *stride = 0; *group = 0;
        if (sys->structured) structured_rng_layout(sys, stride, group);
    });
}
MCG_API int mcg_colour_order(const mcg_system *sys, int32_t *order) {
    return guarded([&] {
        MCG_REQUIRE(sys && order, "NULL argument");
        if (sys->structured) structured_colour_order(sys, order);
        else std::copy(sys->site_of.begin(), sys->site_of.end(), order);
    });
}

MCG_API int mcg_set_params(mcg_system *sys, const double *beta, const double *field) {
    SYS_GUARD({
        if (beta) sys->beta_host.assign(beta, beta + sys->R);
        if (field) sys->field_host.assign(field, field + sys->R);
        MCG_CUDA(cudaMemcpyAsync(sys->d_beta, sys->beta_host.data(), sizeof(double) * sys->R, cudaMemcpyHostToDevice, sys->stream));
        MCG_CUDA(cudaMemcpyAsync(sys->d_field, sys->field_host.data(), sizeof(double) * sys->R, cudaMemcpyHostToDevice, sys->stream));
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
    });
}

// A created system back to the state mcg_create_* leaves it in, with new replica parameters and Philox streams: every
// counter that feeds the RNG (sweep, Wolff step, measurement) restarts at zero, so a recycled system run with (seed, offset)
// produces bit for bit what a fresh system created with them produces - without the allocation, table upload, colouring
// search and module load of a creation.  The spins are whatever the previous job left: callers re-initialise them.
MCG_API int mcg_recycle(mcg_system *sys, const double *beta, const double *field, uint64_t seed, int replica_offset) {
    SYS_GUARD({
        MCG_REQUIRE(!sys->pt && sys->nLabel == sys->R, "a system set up for parallel tempering cannot be recycled");
        MCG_REQUIRE(replica_offset >= 0, "negative replica offset");
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
        sys->seed = seed; sys->replica0 = (uint32_t)replica_offset;
        sys->sweepCtr = 0; sys->wolffCtr = 0; sys->wolffPrimed = false;
        sys->launches = 0; sys->jitLaunches = 0;
        if (sys->d_wmode) {   // frontier/global hybrid bookkeeping: rebuilt on the next Wolff step
            pool_free(sys->d_wmode); pool_free(sys->d_wstamp); pool_free(sys->d_wqueue); pool_free(sys->d_hparent);
            sys->d_wmode = nullptr; sys->d_wstamp = nullptr; sys->d_wqueue = nullptr; sys->d_hparent = nullptr;
        }
        if (beta) sys->beta_host.assign(beta, beta + sys->R);
        if (field) sys->field_host.assign(field, field + sys->R);
        MCG_CUDA(cudaMemcpyAsync(sys->d_beta, sys->beta_host.data(), sizeof(double) * sys->R, cudaMemcpyHostToDevice, sys->stream));
        MCG_CUDA(cudaMemcpyAsync(sys->d_field, sys->field_host.data(), sizeof(double) * sys->R, cudaMemcpyHostToDevice, sys->stream));
        reset_measurements_async(sys);
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
    });
}

MCG_API int mcg_slab_plan(const mcg_lattice_desc *d, int precision, int rank, int world, int32_t *info6) {
    return guarded([&] {
        MCG_REQUIRE(d && info6, "NULL argument");
        structured_slab_plan(d, precision, rank, world, info6);
    });
}
MCG_API int mcg_slab_info(const mcg_system *sys, int32_t *info6) {
    return guarded([&] {
        MCG_REQUIRE(sys && info6, "NULL argument");
        MCG_REQUIRE(sys->structured && structured_is_slab(sys), "not a slab of a decomposed lattice");
        structured_slab_info(sys, info6);
    });
}
MCG_API int mcg_slab_sync(mcg_system *sys) {
    SYS_GUARD({
        MCG_REQUIRE(sys->structured && structured_is_slab(sys), "not a slab of a decomposed lattice");
        structured_slab_exchange_all(sys);
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
    });
}

MCG_API int mcg_init_spins(mcg_system *sys, double flunc) { SYS_GUARD(init_spins(sys, flunc)); }
MCG_API int mcg_set_spins(mcg_system *sys, int replica, const double *spins) { SYS_GUARD(set_spins(sys, replica, spins)); }
MCG_API int mcg_get_spins(mcg_system *sys, int replica, double *spins) { SYS_GUARD(get_spins(sys, replica, spins)); }
MCG_API int mcg_energy(mcg_system *sys, int replica, double *Etot, double *eb, double *eo) { SYS_GUARD(energy(sys, replica, Etot, eb, eo)); }

MCG_API int mcg_metropolis_sweeps(mcg_system *sys, int64_t nsweeps, double pAttempt) {
    SYS_GUARD({ metropolis_sweeps(sys, nsweeps, pAttempt); MCG_CUDA(cudaStreamSynchronize(sys->stream)); });
}
MCG_API int mcg_timed_sweeps(mcg_system *sys, int64_t nsweeps, double pAttempt, int with_measure, double *elapsed_ms) {
    SYS_GUARD({
        MCG_REQUIRE(elapsed_ms, "elapsed_ms is NULL");
        cudaEvent_t e0, e1;
        MCG_CUDA(cudaEventCreate(&e0));
        MCG_CUDA(cudaEventCreate(&e1));
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
        MCG_CUDA(cudaEventRecord(e0, sys->stream));
        if (!with_measure) metropolis_sweeps(sys, nsweeps, pAttempt);
        else
            for (int64_t i = 0; i < nsweeps; i++) measured_sweep(sys, pAttempt);
        MCG_CUDA(cudaEventRecord(e1, sys->stream));
        MCG_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MCG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        MCG_CUDA(cudaGetLastError());
        *elapsed_ms = ms;
    });
}
MCG_API int mcg_wolff_steps(mcg_system *sys, int64_t nsteps) {
    SYS_GUARD({ wolff_steps(sys, nsteps); MCG_CUDA(cudaStreamSynchronize(sys->stream)); });
}
MCG_API int mcg_measure(mcg_system *sys) {
    SYS_GUARD({ measure(sys); MCG_CUDA(cudaStreamSynchronize(sys->stream)); });
}
MCG_API int mcg_reset_measurements(mcg_system *sys) {
    SYS_GUARD({
        reset_measurements_async(sys);
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
    });
}
MCG_API int mcg_results(mcg_system *sys, int replica, double *out, double *groupOut) { SYS_GUARD(results(sys, replica, out, groupOut)); }

MCG_API int mcg_acc_get(mcg_system *sys, int label, double *row, int *n_acc) {
    SYS_GUARD({
        MCG_REQUIRE(label >= 0 && label < sys->nLabel && row, "bad label or NULL row");
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
        MCG_CUDA(cudaMemcpy(row, sys->d_acc + (size_t)label * NACC, sizeof(double) * NACC, cudaMemcpyDeviceToHost));
        if (n_acc) *n_acc = NACC;
    });
}
MCG_API int mcg_acc_set(mcg_system *sys, int label, const double *row) {
    SYS_GUARD({
        MCG_REQUIRE(label >= 0 && label < sys->nLabel && row, "bad label or NULL row");
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
        MCG_CUDA(cudaMemcpy(sys->d_acc + (size_t)label * NACC, row, sizeof(double) * NACC, cudaMemcpyHostToDevice));
    });
}
MCG_API int mcg_wolff_frontier_steps(mcg_system *sys, int replica, int64_t *steps) {
    SYS_GUARD({
        MCG_REQUIRE(replica >= 0 && replica < sys->R && steps, "bad replica index");
        *steps = 0;
        if (sys->d_wmode) {
            int32_t m[WM_N];
            MCG_CUDA(cudaStreamSynchronize(sys->stream));
            MCG_CUDA(cudaMemcpy(m, sys->d_wmode + (size_t)replica * WM_N, sizeof(m), cudaMemcpyDeviceToHost));
            *steps = m[WM_NFRONT];
        }
    });
}

MCG_API int mcg_counters(mcg_system *sys, int replica, int64_t *attempts, int64_t *accepted, int64_t *cluster_sites) {
    SYS_GUARD({
        MCG_REQUIRE(replica >= 0 && replica < sys->R, "bad replica index");
        unsigned long long c[NCNT];
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
        MCG_CUDA(cudaMemcpy(c, sys->d_cnt + (size_t)replica * NCNT, sizeof(c), cudaMemcpyDeviceToHost));
        if (attempts) *attempts = (int64_t)c[CNT_ATTEMPT];
        if (accepted) *accepted = (int64_t)c[CNT_ACCEPT];
        if (cluster_sites) *cluster_sites = (int64_t)c[CNT_CLUSTER];
    });
}

MCG_API int mcg_launch_count(mcg_system *sys, int64_t *launches) {
    return guarded([&] { MCG_REQUIRE(sys && launches, "NULL argument"); *launches = (int64_t)sys->launches; });
}
MCG_API int mcg_jit_launch_count(mcg_system *sys, int64_t *launches) {
    return guarded([&] { MCG_REQUIRE(sys && launches, "NULL argument"); *launches = (int64_t)sys->jitLaunches; });
}
MCG_API int mcg_jit_module_key(mcg_system *sys, int colour, uint64_t *key) {
    return guarded([&] { MCG_REQUIRE(sys && key, "NULL argument"); *key = structured_jit_key(sys, colour); });
}
MCG_API int mcg_profile_passes(mcg_system *sys, int on) {
    return guarded([&] { MCG_REQUIRE(sys, "system is NULL"); sys->profilePasses = on != 0; });
}
MCG_API int mcg_profile_read(mcg_system *sys, double *total_ms, int64_t *nlaunches) {
    SYS_GUARD({
        MCG_REQUIRE(total_ms && nlaunches, "NULL argument");
        MCG_CUDA(cudaStreamSynchronize(sys->stream));
        double tot = 0;
        for (auto &pr : sys->passEvents) {
            float ms = 0;
            MCG_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
            tot += ms;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        *total_ms = tot;
        *nlaunches = (int64_t)sys->passEvents.size();
        sys->passEvents.clear();
    });
}

MCG_API int mcg_run(mcg_system *sys, int algorithm, int64_t nthermal, int64_t nsweep, int64_t ninterval, int spinFrame, double *frames) {
    SYS_GUARD(run(sys, algorithm, nthermal, nsweep, ninterval, spinFrame, frames));
}

static int run_legacy(const mcg_tables *t, int model_expected_ising, int algorithm, int64_t nthermal, int64_t nsweep,
                      int64_t ninterval, double flunc, double h, int spinFrame, uint64_t seed, int precision, double *out,
                      double *frames, double *groupOut) {
    return guarded([&] {
        MCG_REQUIRE(t && out, "NULL argument");
        MCG_REQUIRE((t->model == MCG_ISING) == (model_expected_ising != 0), "wrong entry point for this model");
        mcg_config cfg;
        std::memset(&cfg, 0, sizeof cfg);
        double beta = 1.0;
        cfg.precision = precision; cfg.nReplica = 1; cfg.beta = &beta; cfg.field = &h; cfg.seed = seed; cfg.device = -1;
        std::unique_ptr<mcg_system> sys(create_from_tables(t, &cfg));
        init_spins(sys.get(), flunc);
        run(sys.get(), algorithm, nthermal, nsweep, ninterval, spinFrame, frames);
        results(sys.get(), 0, out, groupOut);
    });
}

MCG_API int mcg_run_on(const mcg_tables *t, int algorithm, int64_t nthermal, int64_t nsweep, int64_t ninterval, double flunc, double h,
               int spinFrame, uint64_t seed, int precision, double out27[27], double *frames, double *groupOut) {
    return run_legacy(t, 0, algorithm, nthermal, nsweep, ninterval, flunc, h, spinFrame, seed, precision, out27, frames, groupOut);
}

MCG_API int mcg_run_ising(const mcg_tables *t, int algorithm, int64_t nthermal, int64_t nsweep, int64_t ninterval, double h, int spinFrame,
                  uint64_t seed, int precision, double out10[10], double *frames) {
    return run_legacy(t, 1, algorithm, nthermal, nsweep, ninterval, 0.0, h, spinFrame, seed, precision, out10, frames, nullptr);
}

}  // extern "C"
