// Host-side description of a resident simulation (tables + replica states in HBM).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "common.cuh"

namespace mcg {

// POD handed to the generic (table-driven) kernels by value.
struct GenArgs {
    int N, maxL, R;
    int nJ, ncls;
    int dupLinks;               // some site pair is linked through more than one slot (forceAdd'ed tables)
    const int32_t *nbrp;        // [maxL][N] neighbour storage positions (pad: self)
    const uint16_t *jtype;      // [maxL][N] index into Jtab (0 = zero tensor)
    const void *Jtab;           // [nJ][9] (Ising [nJ][1]) real
    const uint16_t *cls;        // [N] site class
    const void *clsS;           // [ncls] |S| real
    const void *clsD;           // [ncls][3] real
    const int32_t *site_of;     // [N] storage position -> reference site id
    void *spin;                 // [R][NC][N] real
    const double *beta;         // [R]
    const double *field;        // [R]
    unsigned long long *cnt;    // [R][NCNT]
    RngKey key;
    uint32_t replica0;
};

struct StructuredSystem;  // structured.cu
struct PtState;           // pt.cu: ladder, label holders and the NCCL communicator of an in-library tempering run
void pt_destroy(PtState *p);

}  // namespace mcg

namespace mcg {
// engine.cu internals used by pt.cu
void measured_sweep(mcg_system *s, double pAtt);
void reset_measurements_async(mcg_system *s);
void results_from(mcg_system *s, const double *accBase, const double *gaccBase, int r, double *out, double *groupOut);
}  // namespace mcg

struct mcg_system {
    int model = 0, NC = 0, prec = 64, R = 1;
    int N = 0, maxL = 0;
    bool fullJ = false;
    bool structured = false;
    int device = 0;
    uint64_t seed = 1;
    uint32_t replica0 = 0;
    uint64_t sweepCtr = 0;       // Metropolis sweep counter (Philox counter word)
    uint64_t wolffCtr = 0;       // Wolff step counter
    // colouring
    int C = 0;
    std::vector<int> colourStart;        // [C+1] in storage positions
    int *d_colourStart = nullptr;        // device copy (resident kernel)
    std::vector<int32_t> site_of, pos_of;  // permutation
    std::vector<double> S_host;          // [N] signed S in reference order
    // measurement tables
    int nLat = 0, nTri = 0, nG = 0, maxG = 0, nR = 0, nC = 0;
    int NGlobal = 0, nLatGlobal = 0;     // slab decomposition: sites / cells of the WHOLE lattice (0: this system is the whole lattice)
    int normN() const { return NGlobal ? NGlobal : N; }
    int normLat() const { return nLatGlobal ? nLatGlobal : nLat; }
    bool selfPairs = false;
    bool dupLinks = false;
    // device buffers (generic path)
    int32_t *d_nbrp = nullptr, *d_site_of = nullptr, *d_pos_of = nullptr, *d_pairs = nullptr, *d_tri = nullptr;
    int32_t *d_mi = nullptr, *d_mj = nullptr;
    uint16_t *d_jtype = nullptr, *d_cls = nullptr;
    void *d_Jtab = nullptr, *d_clsS = nullptr, *d_clsD = nullptr, *d_spin = nullptr;
    double *d_signS = nullptr;           // [N] signed S (storage order) for init
    int nJ = 0, ncls = 0;
    double *d_beta = nullptr, *d_field = nullptr, *d_sums = nullptr, *d_acc = nullptr;
    // parallel tempering: accumulators live per temperature LABEL (nLabel >= R slots), replica r
    // accumulates into slot d_slot[r]; d_last[r] = {reduced energy, Mx, My, Mz} of the last measured sweep
    int nLabel = 0;
    int32_t *d_slot = nullptr;
    double *d_last = nullptr;
    std::vector<int32_t> slot_host;
    unsigned long long *d_cnt = nullptr;
    double *d_scratch = nullptr;         // [2N] per-site energies / frame staging (3N)
    // block-spin (RG) and orbital-group statistics (table path)
    int32_t *d_rPos = nullptr, *d_rCl = nullptr, *d_rNbrRow = nullptr, *d_rNl = nullptr, *d_pairRowI = nullptr, *d_pairRowJ = nullptr;
    int32_t *d_groups = nullptr;
    double *d_ms = nullptr, *d_rsums = nullptr, *d_gsum = nullptr, *d_gacc = nullptr;
    double rg_ci = 0, rg_cj = 0, rg_cij = 0;
    uint64_t measCtr = 0;
    // wolff
    int32_t *d_parent = nullptr;         // [2][R][N] double-buffered union-find forest
    void *d_proj = nullptr;              // [2][R][N] real
    bool wolffPrimed = false;            // buffers of the next step were prepared by the previous step's flip kernel
    bool isoNoOnsite = false;            // every J is a multiple of the identity and every D is 0 (no exchange-anisotropy residual)
    double *d_wres = nullptr;            // [R][2] residual, cluster size
    // frontier/global hybrid (kernels_wolff.cuh: k_wolff_frontier): per-replica state words, visit stamps, member queues and
    // its own forest pair (the plain sequence and the cooperative kernel keep d_parent to themselves)
    int32_t *d_wmode = nullptr, *d_wqueue = nullptr, *d_hparent = nullptr;
    uint32_t *d_wstamp = nullptr;
    int wolffCap = 0;
    uint32_t wolffTag = 0;
    std::vector<double> beta_host, field_host;
    cudaStream_t stream = nullptr;
    // instrumentation: kernels launched so far; optional CUDA-event timing of the colour-pass kernel
    uint64_t launches = 0;
    uint64_t jitLaunches = 0;     // of those, launches of NVRTC-specialised kernels (mcg_pass_m0/m1, mcg_topo)
    bool profilePasses = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> passEvents;
    mcg::StructuredSystem *st = nullptr;   // structured (descriptor) path state, owned
    mcg::PtState *pt = nullptr;            // parallel-tempering state (mcg_pt_setup), owned
    size_t real_size() const { return prec == 8 ? 1 : (prec == 32 ? 4 : 8); }   // prec 8: Ising spins as int8 (structured path)
    ~mcg_system();
};
