/*
 * hb2_level.cu -- device-resident patch level: the C side of "seam 2" (SURVEY.md 8b).
 *
 * RungeKuttaLevelIntegrator::advanceLevel (src/algs/integrator/RungeKuttaLevelIntegrator.cpp:1457-1929) walks the patches of
 * one level per RK stage: xfer::RefineSchedule::fillData (:1568, :1701), then RungeKuttaPatchStrategy::
 * computeFluxesAndSourcesOnPatch + advanceSingleStepOnPatch per patch (:1724-1739).  With the patch data resident in HBM
 * the same loop runs here for ALL patches a rank owns on one level (boxes of any sizes, in level index space): the
 * conservative variables of every patch are registered once (three state buffers per patch: U^n and the two intermediate
 * states of SSP-RK3), the same-level ghost fill between the rank's patches -- periodic images included -- is ONE kernel
 * launch over a descriptor table built at registration, and every patch advances with the fused stage of its shape's plan
 * (hb2_fused_stage_dev).  Host memory is touched only by hb2_level_upload_patch / hb2_level_download_patch.
 * The entry points are declared in include/hamers_b200.h; the C++ class RungeKuttaPatchStrategyB200
 * (hamers_b200/host) drives them with the reference's method names.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "hamers_b200.h"

namespace hb2 {
int set_error(int code, const std::string& msg);
}
using hb2::set_error;

#define HB2L_CUDA(call)                                                                           \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return set_error(-100 - (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

namespace {

struct CopyDesc {
    int dst, src;              /* patch indices */
    int ext[3];
    long long dst_off, src_off; /* ghost-box index of the region's first cell in the destination / source patch */
    long long dst_cs[2], src_cs[2];
};

struct LevelPatch {
    int lo[3], n[3];
    int shape;
    long long ncell_g;
    double* S[3];              /* three state buffers, each ncomp * ncell_g doubles */
};

/* grid: (x, descriptor); one descriptor = one box-to-box copy of every component.  32-bit index arithmetic (a box has fewer
 * than 2^31 cells), components in the outer loop; gridDim.x is sized by the largest box of the launch (fill_blocks): with a
 * fixed 8 blocks per descriptor the z-ghost planes of a 512 x 512 slab took 2.2 ms per launch -- more than the four kernels of
 * the slab's stage together. */
__global__ void __launch_bounds__(256) k_level_fill(const CopyDesc* __restrict__ D, double* const* __restrict__ ptrs, int ncomp)
{
    const CopyDesc d = D[blockIdx.y];
    const unsigned e0 = (unsigned)d.ext[0], e1 = (unsigned)d.ext[1];
    const unsigned nb = e0 * e1 * (unsigned)d.ext[2];
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < nb; r += gridDim.x * blockDim.x) {
        const unsigned q = r / e0, i = r - q * e0;
        const unsigned k = q / e1, j = q - k * e1;
        const long long xd = d.dst_off + i + d.dst_cs[0] * j + d.dst_cs[1] * k;
        const long long xs = d.src_off + i + d.src_cs[0] * j + d.src_cs[1] * k;
        for (int c = 0; c < ncomp; c++) ptrs[d.dst * ncomp + c][xd] = ptrs[d.src * ncomp + c][xs];
    }
}

/* blocks along x for a launch whose largest box has max_cells cells: 8 cells per thread, between 8 and 1024 blocks */
inline unsigned fill_blocks(long long max_cells)
{
    long long b = (max_cells + 256 * 8 - 1) / (256 * 8);
    return (unsigned)(b < 8 ? 8 : (b > 1024 ? 1024 : b));
}

}  // namespace

struct hb2_level_s {
    hb2_patch_desc model;
    int dim, ncomp, neq, g, device;
    int level_n[3], periodic_mask;
    std::vector<LevelPatch> patches;
    std::vector<hb2_plan_t> plans;              /* one per distinct patch shape */
    std::vector<std::vector<int>> shapes;
    int where[3];                               /* buffer index of U^(0), U^(1), U^(2) of the running step; where[0] = current */
    CopyDesc* d_desc;
    int ndesc;
    double** d_ptrs[3];                         /* device pointer tables [patch * ncomp + c], one per buffer */
    cudaStream_t stream;
    long long launches;
    double* d_sr;                               /* 4 doubles per patch: spectral radii */
    /* pipelined host-memory advance (hb2_level_advance_host) */
    std::vector<int> desc_begin;                /* descriptors of destination patch p: [desc_begin[p], desc_begin[p + 1]) */
    std::vector<long long> desc_max_cells;      /* largest box among the descriptors of destination patch p; [npatch] = level-wide */
    std::vector<std::vector<int>> sources;      /* patches whose data the ghost fill of patch p reads */
    cudaStream_t copy_in, copy_out;
    std::vector<cudaEvent_t> ev_up, ev_done, ev_lo;   /* ev_lo[p]: the first g interior planes of slab p are on the device */
    /* second compute stream with its own plans (a plan owns the running right-hand side and the sensor bytes of the stage in
     * flight): tasks of even / odd patches alternate between the two, so that the tails of the small grids of thin slabs overlap */
    cudaStream_t stream2;
    std::vector<hb2_plan_t> plans2;
    std::vector<cudaEvent_t> ev_task;           /* [stage * npatch + patch] */
    /* HB2_LEVEL_TRACE=1: timed events of the pipelined host advance (start, upload, every task, download per patch), printed
     * to stderr after the step */
    std::vector<cudaEvent_t> tr;
};

namespace {

int build_fill_table(hb2_level_t L)
{
    const int dim = L->dim, g = L->g;
    std::vector<CopyDesc> D;
    const int np = (int)L->patches.size();
    for (int p = 0; p < np; p++) {
        const LevelPatch& P = L->patches[p];
        int shift[3];
        for (shift[2] = (dim == 3 ? -1 : 0); shift[2] <= (dim == 3 ? 1 : 0); shift[2]++)
            for (shift[1] = -1; shift[1] <= 1; shift[1]++)
                for (shift[0] = -1; shift[0] <= 1; shift[0]++) {
                    bool ok = true;
                    for (int a = 0; a < dim; a++)
                        if (shift[a] != 0 && !((L->periodic_mask >> a) & 1)) ok = false;
                    if (!ok) continue;
                    for (int q = 0; q < np; q++) {
                        const LevelPatch& Q = L->patches[q];
                        if (q == p && !shift[0] && !shift[1] && !shift[2]) continue;
                        CopyDesc d;
                        memset(&d, 0, sizeof d);
                        bool empty = false;
                        int rlo[3] = {0, 0, 0};
                        for (int a = 0; a < 3; a++) {
                            if (a >= dim) {
                                d.ext[a] = 1;
                                continue;
                            }
                            const int qlo = Q.lo[a] + shift[a] * L->level_n[a], qhi = qlo + Q.n[a];
                            const int glo = P.lo[a] - g, ghi = P.lo[a] + P.n[a] + g;
                            const int lo = qlo > glo ? qlo : glo, hi = qhi < ghi ? qhi : ghi;
                            if (hi <= lo) empty = true;
                            rlo[a] = lo;
                            d.ext[a] = hi - lo;
                        }
                        if (empty) continue;
                        d.dst = p;
                        d.src = q;
                        const long long pcs1 = P.n[0] + 2 * g, pcs2 = pcs1 * (P.n[1] + 2 * g);
                        const long long qcs1 = Q.n[0] + 2 * g, qcs2 = qcs1 * (Q.n[1] + 2 * g);
                        d.dst_cs[0] = pcs1;
                        d.dst_cs[1] = dim == 3 ? pcs2 : 0;
                        d.src_cs[0] = qcs1;
                        d.src_cs[1] = dim == 3 ? qcs2 : 0;
                        d.dst_off = (rlo[0] - P.lo[0] + g) + pcs1 * (rlo[1] - P.lo[1] + g) + (dim == 3 ? pcs2 * (rlo[2] - P.lo[2] + g) : 0);
                        const int s0 = rlo[0] - shift[0] * L->level_n[0] - Q.lo[0], s1 = rlo[1] - shift[1] * L->level_n[1] - Q.lo[1];
                        const int s2 = dim == 3 ? rlo[2] - shift[2] * L->level_n[2] - Q.lo[2] : 0;
                        d.src_off = (s0 + g) + qcs1 * (s1 + g) + (dim == 3 ? qcs2 * (s2 + g) : 0);
                        D.push_back(d);
                    }
                }
    }
    /* descriptors were generated destination by destination: remember the ranges and who feeds whom */
    L->desc_begin.assign(np + 1, 0);
    L->sources.assign(np, std::vector<int>());
    for (size_t k = 0; k < D.size(); k++) {
        L->desc_begin[D[k].dst + 1] = (int)k + 1;
        bool seen = false;
        for (int q : L->sources[D[k].dst]) seen = seen || q == D[k].src;
        if (!seen) L->sources[D[k].dst].push_back(D[k].src);
    }
    for (int p = 0; p < np; p++)
        if (L->desc_begin[p + 1] < L->desc_begin[p]) L->desc_begin[p + 1] = L->desc_begin[p];
    L->desc_max_cells.assign(np + 1, 0);
    for (size_t k = 0; k < D.size(); k++) {
        const long long cells = (long long)D[k].ext[0] * D[k].ext[1] * D[k].ext[2];
        if (cells > L->desc_max_cells[D[k].dst]) L->desc_max_cells[D[k].dst] = cells;
        if (cells > L->desc_max_cells[np]) L->desc_max_cells[np] = cells;
    }
    L->ndesc = (int)D.size();
    if (L->ndesc) {
        HB2L_CUDA(cudaMalloc(&L->d_desc, sizeof(CopyDesc) * D.size()));
        HB2L_CUDA(cudaMemcpy(L->d_desc, D.data(), sizeof(CopyDesc) * D.size(), cudaMemcpyHostToDevice));
    }
    return 0;
}

}  // namespace

namespace {
/* what travels between host and device for one component of a patch: the planes (rows in 2-D) of the slowest direction
 * that hold interior cells -- one contiguous range of the ghost-box array; the ghost planes beyond are filled on the device */
void transfer_range(hb2_level_t L, const LevelPatch& P, long long* off, long long* cnt)
{
    long long plane = 1;
    for (int a = 0; a < L->dim - 1; a++) plane *= P.n[a] + 2 * L->g;
    *off = plane * L->g;
    *cnt = plane * P.n[L->dim - 1];
}
}  // namespace

extern "C" {

int hb2_level_create(const hb2_patch_desc* model, int32_t npatch, const int32_t* lo, const int32_t* hi, const int32_t level_n[3],
                     int32_t periodic_mask, hb2_level_t* out)
{
    if (!model || !lo || !hi || !level_n || !out) return set_error(-1, "hb2_level_create: null argument");
    if (npatch < 1) return set_error(-40, "hb2_level_create: a level needs at least one patch");
    int32_t neq = 0, ncomp = 0;
    hb2_patch_desc probe = *model;
    for (int a = 0; a < 3; a++) probe.n[a] = 1;
    if (hb2_num_eqn(&probe, &neq) || hb2_num_comp(&probe, &ncomp)) return -1;
    hb2_level_t L = new hb2_level_s();
    L->model = *model;
    L->dim = model->dim;
    L->neq = neq;
    L->ncomp = ncomp;
    L->g = model->num_ghosts > 0 ? model->num_ghosts : HB2_GHOSTS;
    L->periodic_mask = periodic_mask;
    L->d_desc = nullptr;
    L->d_sr = nullptr;
    L->launches = 0;
    L->copy_in = L->copy_out = nullptr;
    L->stream2 = nullptr;
    for (int b = 0; b < 3; b++) {
        L->where[b] = b;
        L->d_ptrs[b] = nullptr;
    }
    for (int a = 0; a < 3; a++) L->level_n[a] = a < L->dim ? level_n[a] : 1;
    for (int a = 0; a < L->dim; a++)
        if (((periodic_mask >> a) & 1) && L->level_n[a] < L->g) {
            delete L;
            return set_error(-41, "hb2_level_create: a periodic direction must be at least one ghost width long");
        }
    L->device = model->device;
    if (L->device < 0) HB2L_CUDA(cudaGetDevice(&L->device));
    HB2L_CUDA(cudaSetDevice(L->device));
    HB2L_CUDA(cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking));
    HB2L_CUDA(cudaStreamCreateWithFlags(&L->copy_in, cudaStreamNonBlocking));
    HB2L_CUDA(cudaStreamCreateWithFlags(&L->copy_out, cudaStreamNonBlocking));
    std::map<std::vector<int>, int> shape_of;
    for (int p = 0; p < npatch; p++) {
        LevelPatch P;
        memset(&P, 0, sizeof P);
        std::vector<int> shp(3, 1);
        P.ncell_g = 1;
        for (int a = 0; a < 3; a++) {
            P.lo[a] = a < L->dim ? lo[3 * p + a] : 0;
            P.n[a] = a < L->dim ? hi[3 * p + a] - lo[3 * p + a] : 1;
            if (P.n[a] < 1) {
                hb2_level_destroy(L);
                return set_error(-3, "hb2_level_create: empty patch box");
            }
            shp[a] = P.n[a];
            P.ncell_g *= (a < L->dim) ? P.n[a] + 2 * L->g : 1;
        }
        auto it = shape_of.find(shp);
        if (it == shape_of.end()) {
            hb2_patch_desc d = *model;
            for (int a = 0; a < 3; a++) d.n[a] = shp[a];
            d.device = L->device;
            hb2_plan_t plan = nullptr;
            int rc = hb2_plan_create(&d, &plan);
            if (rc) {
                hb2_level_destroy(L);
                return rc;
            }
            hb2_plan_set_stream(plan, (void*)L->stream);
            L->plans.push_back(plan);
            L->shapes.push_back(shp);
            it = shape_of.insert(std::make_pair(shp, (int)L->plans.size() - 1)).first;
        }
        P.shape = it->second;
        for (int b = 0; b < 3; b++) {
            HB2L_CUDA(cudaMalloc(&P.S[b], sizeof(double) * (size_t)P.ncell_g * L->ncomp));
            HB2L_CUDA(cudaMemsetAsync(P.S[b], 0, sizeof(double) * (size_t)P.ncell_g * L->ncomp, L->stream));
        }
        L->patches.push_back(P);
    }
    for (int b = 0; b < 3; b++) {
        std::vector<double*> tab((size_t)npatch * L->ncomp);
        for (int p = 0; p < npatch; p++)
            for (int c = 0; c < L->ncomp; c++) tab[(size_t)p * L->ncomp + c] = L->patches[p].S[b] + (size_t)c * L->patches[p].ncell_g;
        HB2L_CUDA(cudaMalloc(&L->d_ptrs[b], sizeof(double*) * tab.size()));
        HB2L_CUDA(cudaMemcpy(L->d_ptrs[b], tab.data(), sizeof(double*) * tab.size(), cudaMemcpyHostToDevice));
    }
    HB2L_CUDA(cudaMalloc(&L->d_sr, sizeof(double) * 4 * (size_t)npatch));
    L->ev_up.resize(npatch);
    L->ev_lo.resize(npatch);
    L->ev_done.resize(npatch);
    for (int p = 0; p < npatch; p++) {
        HB2L_CUDA(cudaEventCreateWithFlags(&L->ev_up[p], cudaEventDisableTiming));
        HB2L_CUDA(cudaEventCreateWithFlags(&L->ev_lo[p], cudaEventDisableTiming));
        HB2L_CUDA(cudaEventCreateWithFlags(&L->ev_done[p], cudaEventDisableTiming));
    }
    int rc = build_fill_table(L);
    if (rc) {
        hb2_level_destroy(L);
        return rc;
    }
    *out = L;
    return 0;
}

int hb2_level_destroy(hb2_level_t L)
{
    if (!L) return 0;
    cudaSetDevice(L->device);
    if (L->stream) cudaStreamSynchronize(L->stream);
    for (auto& P : L->patches)
        for (int b = 0; b < 3; b++) cudaFree(P.S[b]);
    for (auto plan : L->plans) hb2_plan_destroy(plan);
    for (int b = 0; b < 3; b++) cudaFree(L->d_ptrs[b]);
    cudaFree(L->d_desc);
    cudaFree(L->d_sr);
    for (auto e : L->ev_task) cudaEventDestroy(e);
    for (auto plan : L->plans2) hb2_plan_destroy(plan);
    if (L->stream2) cudaStreamDestroy(L->stream2);
    for (auto e : L->ev_up) cudaEventDestroy(e);
    for (auto e : L->ev_lo) cudaEventDestroy(e);
    for (auto e : L->ev_done) cudaEventDestroy(e);
    if (L->copy_in) cudaStreamDestroy(L->copy_in);
    if (L->copy_out) cudaStreamDestroy(L->copy_out);
    if (L->stream) cudaStreamDestroy(L->stream);
    delete L;
    return 0;
}

int hb2_level_num_patches(hb2_level_t L) { return L ? (int)L->patches.size() : -1; }

int64_t hb2_level_launch_count(hb2_level_t L)
{
    if (!L) return -1;
    long long n = L->launches;
    for (auto plan : L->plans) n += hb2_plan_launch_count(plan);
    for (auto plan : L->plans2) n += hb2_plan_launch_count(plan);
    return n;
}

int hb2_level_synchronize(hb2_level_t L)
{
    if (!L) return set_error(-1, "null level");
    HB2L_CUDA(cudaSetDevice(L->device));
    HB2L_CUDA(cudaStreamSynchronize(L->stream));
    return 0;
}

/* host -> device: the INTERIOR of the patch's current state from ghost-box host arrays (SAMRAI CellData pointers) */
int hb2_level_upload_patch(hb2_level_t L, int32_t patch, const double* const* U_host)
{
    if (!L || !U_host) return set_error(-1, "hb2_level_upload_patch: null argument");
    if (patch < 0 || patch >= (int)L->patches.size()) return set_error(-42, "patch index out of range");
    HB2L_CUDA(cudaSetDevice(L->device));
    const LevelPatch& P = L->patches[patch];
    long long off, cnt;
    transfer_range(L, P, &off, &cnt);
    for (int c = 0; c < L->ncomp; c++)
        HB2L_CUDA(cudaMemcpyAsync(P.S[L->where[0]] + (size_t)c * P.ncell_g + off, U_host[c] + off, sizeof(double) * (size_t)cnt,
                                  cudaMemcpyHostToDevice, L->stream));
    return 0;
}

int hb2_level_download_patch(hb2_level_t L, int32_t patch, double* const* U_host)
{
    if (!L || !U_host) return set_error(-1, "hb2_level_download_patch: null argument");
    if (patch < 0 || patch >= (int)L->patches.size()) return set_error(-42, "patch index out of range");
    HB2L_CUDA(cudaSetDevice(L->device));
    const LevelPatch& P = L->patches[patch];
    long long off, cnt;
    transfer_range(L, P, &off, &cnt);
    for (int c = 0; c < L->ncomp; c++)
        HB2L_CUDA(cudaMemcpyAsync(U_host[c] + off, P.S[L->where[0]] + (size_t)c * P.ncell_g + off, sizeof(double) * (size_t)cnt,
                                  cudaMemcpyDeviceToHost, L->stream));
    HB2L_CUDA(cudaStreamSynchronize(L->stream));
    return 0;
}

int hb2_level_patch_state_dev(hb2_level_t L, int32_t patch, int32_t state, double** U_dev)
{
    if (!L || !U_dev) return set_error(-1, "hb2_level_patch_state_dev: null argument");
    if (patch < 0 || patch >= (int)L->patches.size() || state < 0 || state > 2) return set_error(-42, "patch / state index out of range");
    const LevelPatch& P = L->patches[patch];
    for (int c = 0; c < L->ncomp; c++) U_dev[c] = P.S[L->where[state]] + (size_t)c * P.ncell_g;
    return 0;
}

/* same-level ghost fill of intermediate state `state` (0 = current) of every patch, periodic images included */
int hb2_level_fill_ghosts(hb2_level_t L, int32_t state)
{
    if (!L) return set_error(-1, "null level");
    if (state < 0 || state > 2) return set_error(-42, "state index out of range");
    if (!L->ndesc) return 0;
    HB2L_CUDA(cudaSetDevice(L->device));
    dim3 grid(fill_blocks(L->desc_max_cells[L->patches.size()]), (unsigned)L->ndesc);
    k_level_fill<<<grid, 256, 0, L->stream>>>(L->d_desc, L->d_ptrs[L->where[state]], L->ncomp);
    L->launches++;
    HB2L_CUDA(cudaGetLastError());
    return 0;
}

namespace {
/* output buffer of stage ncoef - 1: U^(ncoef) takes buffer where[ncoef] for ncoef < 3; the last stage of a three-stage
 * scheme overwrites the buffer of the older state whose alpha is zero (SSP-RK3: U^(1)) */
int stage_output_buffer(hb2_level_t L, int ncoef, const double* alpha)
{
    if (ncoef < 3) return L->where[ncoef];
    int out = -1;
    for (int m = 0; m < ncoef - 1; m++)
        if (alpha[m] == 0.0) out = L->where[m];
    return out;
}
}  // namespace

/* computeFluxesAndSourcesOnPatch + advanceSingleStepOnPatch of ONE patch for RK stage ncoef - 1 (fused; the flux is taken of
 * U^(ncoef - 1), whose ghosts must be filled: hb2_level_fill_ghosts(L, ncoef - 1)) */
static int advance_stage_patch_on(hb2_level_t L, int patch, int ncoef, const double* alpha, const double* beta, double dt, hb2_plan_t plan)
{
    const int out = stage_output_buffer(L, ncoef, alpha);
    if (out < 0) return set_error(-43, "a third stage needs one alpha == 0 among the older states (three buffers per patch)");
    const LevelPatch& P = L->patches[patch];
    const double* Uint[3 * HB2_MAX_COMP];
    double* Uout[HB2_MAX_COMP];
    for (int m = 0; m < ncoef; m++)
        for (int c = 0; c < L->ncomp; c++) Uint[m * L->ncomp + c] = P.S[L->where[m]] + (size_t)c * P.ncell_g;
    for (int c = 0; c < L->ncomp; c++) Uout[c] = P.S[out] + (size_t)c * P.ncell_g;
    return hb2_fused_stage_dev(plan, ncoef, alpha, beta, Uint, dt, Uout);
}

int hb2_level_advance_stage_patch(hb2_level_t L, int32_t patch, int32_t ncoef, const double* alpha, const double* beta, double dt)
{
    if (!L || !alpha || !beta) return set_error(-1, "hb2_level_advance_stage_patch: null argument");
    if (ncoef < 1 || ncoef > 3) return set_error(-15, "hb2_level_advance_stage_patch: ncoef must be 1..3 (SSP-RK3)");
    if (patch < 0 || patch >= (int)L->patches.size()) return set_error(-42, "patch index out of range");
    HB2L_CUDA(cudaSetDevice(L->device));
    const int out = stage_output_buffer(L, ncoef, alpha);
    if (out < 0) return set_error(-43, "a third stage needs one alpha == 0 among the older states (three buffers per patch)");
    const LevelPatch& P = L->patches[patch];
    const double* Uint[3 * HB2_MAX_COMP];
    double* Uout[HB2_MAX_COMP];
    for (int m = 0; m < ncoef; m++)
        for (int c = 0; c < L->ncomp; c++) Uint[m * L->ncomp + c] = P.S[L->where[m]] + (size_t)c * P.ncell_g;
    for (int c = 0; c < L->ncomp; c++) Uout[c] = P.S[out] + (size_t)c * P.ncell_g;
    return hb2_fused_stage_dev(L->plans[P.shape], ncoef, alpha, beta, Uint, dt, Uout);
}

/* after every patch has advanced through stage ncoef - 1: U^(ncoef) is where stage_output_buffer says; after the last stage it
 * becomes the current state and the other two buffers are free */
int hb2_level_end_stage(hb2_level_t L, int32_t ncoef, const double* alpha, int32_t last_stage)
{
    if (!L || !alpha) return set_error(-1, "hb2_level_end_stage: null argument");
    const int out = stage_output_buffer(L, ncoef, alpha);
    if (out < 0) return set_error(-43, "a third stage needs one alpha == 0 among the older states (three buffers per patch)");
    if (last_stage) {
        int rest[2], k = 0;
        for (int b = 0; b < 3; b++)
            if (b != out) rest[k++] = b;
        L->where[0] = out;
        L->where[1] = rest[0];
        L->where[2] = rest[1];
    }
    return 0;
}

int hb2_level_advance_stage(hb2_level_t L, int32_t ncoef, const double* alpha, const double* beta, double dt, int32_t last_stage)
{
    if (!L) return set_error(-1, "null level");
    for (int pi = 0; pi < (int)L->patches.size(); pi++) {
        int rc = hb2_level_advance_stage_patch(L, pi, ncoef, alpha, beta, dt);
        if (rc) return rc;
    }
    return hb2_level_end_stage(L, ncoef, alpha, last_stage);
}

/* RungeKuttaLevelIntegrator::advanceLevel for the whole level: per stage the ghost fill, then every patch (alpha / beta:
 * row-major [nstages][nstages]) */
int hb2_level_advance(hb2_level_t L, int32_t nstages, const double* alpha, const double* beta, double dt)
{
    if (!L || !alpha || !beta) return set_error(-1, "hb2_level_advance: null argument");
    if (nstages < 1 || nstages > 3) return set_error(-20, "hb2_level_advance supports 1..3 stages");
    for (int sn = 0; sn < nstages; sn++) {
        int rc = hb2_level_fill_ghosts(L, sn);
        if (rc) return rc;
        rc = hb2_level_advance_stage(L, sn + 1, alpha + sn * nstages, beta + sn * nstages, dt, sn == nstages - 1);
        if (rc) return rc;
    }
    return 0;
}

/* advanceLevel on HOST memory, pipelined over the patches.  U_host[p * num_comp + c]: ghost-box array of component c of
 * patch p (pinned memory gives asynchronous copies); on return it holds the new state (interior planes of the slowest
 * direction are transferred, like hb2_level_upload_patch / _download_patch).
 * A time step on host memory is PCIe-bound (the whole state goes in and comes out), but PCIe is full duplex and a patch
 * depends on its neighbours only: stage s of patch p needs stage s - 1 of p and of the patches its ghost cells come from.
 * So the patches are uploaded on one copy stream, every (patch, stage) task is enqueued on the compute stream in the order in
 * which the uploads make it runnable (a wavefront that trails the upload by a few patches), and a patch is downloaded on a
 * second copy stream as soon as its last stage is done -- the download of the early patches runs while the late ones are
 * still being uploaded.  Buffers: stage outputs rotate like in hb2_level_advance; a patch overwrites U^(1) in its last stage
 * only after every neighbour's second stage (which read it) is done, because its own last stage depends on those. */
int hb2_level_advance_host(hb2_level_t L, int32_t nstages, const double* alpha, const double* beta, double dt,
                           double* const* U_host)
{
    if (!L || !alpha || !beta || !U_host) return set_error(-1, "hb2_level_advance_host: null argument");
    if (nstages < 1 || nstages > 3) return set_error(-20, "hb2_level_advance_host supports 1..3 stages");
    HB2L_CUDA(cudaSetDevice(L->device));
    const int np = (int)L->patches.size();
    static const bool trace = [] {
        const char* v = getenv("HB2_LEVEL_TRACE");
        return v && *v && atoi(v) != 0;
    }();
    /* tr[0] start; tr[1 + p] upload of p done; tr[1 + np + sn * np + p] task done; tr[1 + 4 np + p] download of p done;
     * tr[1 + 5 np + sn * np + p] task started (behind its stream's waits) */
    if (trace && L->tr.empty()) {
        L->tr.resize((size_t)1 + 8 * np);
        for (auto& e : L->tr) HB2L_CUDA(cudaEventCreate(&e));
    }
    if (trace) HB2L_CUDA(cudaEventRecord(L->tr[0], L->copy_in));
    /* When every patch spans the level in all but the slowest direction (a stack of slabs), a patch's ghost cells come from
     * the FIRST and LAST g interior planes of its neighbours only. */
    bool slabs = np > 1;
    for (int p = 0; p < np && slabs; p++)
        for (int a = 0; a < L->dim - 1; a++) slabs = slabs && L->patches[p].n[a] == L->level_n[a];
    for (int p = 0; p < np && slabs; p++) slabs = L->patches[p].n[L->dim - 1] > 2 * L->g;
    /* HB2_LEVEL_EDGES=all: the edge planes of EVERY slab go up before any bulk (the first form of this schedule: with 18 slabs
     * that is 28 % of the data in front of the first stage).  Default: the slabs go up whole, in the order of the slowest
     * direction, low edge first; the first stage of a slab waits for its own planes and for the low edge of the next slab
     * (the very next copy); only the high edge of the LAST slab, which the first slab needs across the periodic seam, goes
     * up ahead of everything. */
    static const bool edges_all = [] {
        const char* v = getenv("HB2_LEVEL_EDGES");
        return v && std::string(v) == "all";
    }();
    const int zd = L->dim - 1;
    std::vector<int> uorder, pos(np), znext(np, -1);
    bool seam = false;
    if (slabs && !edges_all) {
        for (int p = 0; p < np; p++) uorder.push_back(p);
        for (int i = 1; i < np; i++)
            for (int j = i; j > 0 && L->patches[uorder[j]].lo[zd] < L->patches[uorder[j - 1]].lo[zd]; j--) std::swap(uorder[j], uorder[j - 1]);
        for (int i = 0; i + 1 < np; i++) znext[uorder[i]] = uorder[i + 1];
        const LevelPatch &first = L->patches[uorder[0]], &last = L->patches[uorder[np - 1]];
        seam = ((L->periodic_mask >> zd) & 1) && first.lo[zd] == 0 && last.lo[zd] + last.n[zd] == L->level_n[zd];
        if (seam) znext[uorder[np - 1]] = uorder[0];
    } else {
        /* index order, except that a last patch that feeds the first one (periodic stack of slabs) goes first */
        bool wrap = false;
        for (int q : L->sources[0]) wrap = wrap || (q == np - 1 && np > 2);
        if (wrap) uorder.push_back(np - 1);
        for (int p = 0; p < np - (wrap ? 1 : 0); p++) uorder.push_back(p);
    }
    for (int i = 0; i < np; i++) pos[uorder[i]] = i;
    auto copy_planes = [&](int p, long long first_plane_off, long long count) -> int {
        const LevelPatch& P = L->patches[p];
        for (int c = 0; c < L->ncomp; c++)
            HB2L_CUDA(cudaMemcpyAsync(P.S[L->where[0]] + (size_t)c * P.ncell_g + first_plane_off, U_host[(size_t)p * L->ncomp + c] + first_plane_off,
                                      sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, L->copy_in));
        return 0;
    };
    if (slabs && !edges_all) {
        if (seam) {
            const int p = uorder[np - 1];
            long long off, cnt;
            transfer_range(L, L->patches[p], &off, &cnt);
            const long long edge = cnt / L->patches[p].n[zd] * L->g;
            if (copy_planes(p, off + cnt - edge, edge)) return -1;
        }
        HB2L_CUDA(cudaEventRecord(L->ev_done[0], L->copy_in));        /* borrowed as "the seam planes are on the device" */
        HB2L_CUDA(cudaStreamWaitEvent(L->stream, L->ev_done[0], 0));
        for (int i = 0; i < np; i++) {
            const int p = uorder[i];
            long long off, cnt;
            transfer_range(L, L->patches[p], &off, &cnt);
            const long long edge = cnt / L->patches[p].n[zd] * L->g;
            if (copy_planes(p, off, edge)) return -1;
            HB2L_CUDA(cudaEventRecord(L->ev_lo[p], L->copy_in));
            const long long rest = cnt - edge - ((seam && i == np - 1) ? edge : 0);
            if (copy_planes(p, off + edge, rest)) return -1;
            HB2L_CUDA(cudaEventRecord(L->ev_up[p], L->copy_in));
            if (trace) HB2L_CUDA(cudaEventRecord(L->tr[1 + p], L->copy_in));
        }
    } else {
        if (slabs) {
            for (int p = 0; p < np; p++) {
                long long off, cnt;
                transfer_range(L, L->patches[p], &off, &cnt);
                const long long edge = cnt / L->patches[p].n[zd] * L->g;
                if (copy_planes(p, off, edge) || copy_planes(p, off + cnt - edge, edge)) return -1;
            }
            HB2L_CUDA(cudaEventRecord(L->ev_done[0], L->copy_in));        /* borrowed as "all edge planes are on the device" */
            HB2L_CUDA(cudaStreamWaitEvent(L->stream, L->ev_done[0], 0));
        }
        for (int i = 0; i < np; i++) {
            const int p = uorder[i];
            long long off, cnt;
            transfer_range(L, L->patches[p], &off, &cnt);
            if (slabs) {
                const long long edge = cnt / L->patches[p].n[zd] * L->g;
                off += edge;
                cnt -= 2 * edge;
            }
            if (copy_planes(p, off, cnt)) return -1;
            HB2L_CUDA(cudaEventRecord(L->ev_up[p], L->copy_in));
            if (trace) HB2L_CUDA(cudaEventRecord(L->tr[1 + p], L->copy_in));
        }
    }
    /* readiness of every (patch, stage) task in units of upload positions, then the enqueue order */
    std::vector<std::vector<int>> ready(nstages, std::vector<int>(np, 0));
    for (int sn = 0; sn < nstages; sn++)
        for (int p = 0; p < np; p++) {
            int r = sn == 0 ? pos[p] : ready[sn - 1][p];
            if (sn > 0 || !slabs)
                for (int q : L->sources[p]) {
                    const int rq = sn == 0 ? pos[q] : ready[sn - 1][q];
                    r = rq > r ? rq : r;
                }
            ready[sn][p] = r;
        }
    struct Task {
        int r, sn, p;
    };
    std::vector<Task> tasks;
    for (int sn = 0; sn < nstages; sn++)
        for (int p = 0; p < np; p++) tasks.push_back(Task{ready[sn][p], sn, p});
    for (size_t i = 1; i < tasks.size(); i++)
        for (size_t j = i; j > 0; j--) {
            const Task &x = tasks[j], &y = tasks[j - 1];
            const bool less = x.r < y.r || (x.r == y.r && (x.sn < y.sn || (x.sn == y.sn && x.p < y.p)));
            if (!less) break;
            std::swap(tasks[j], tasks[j - 1]);
        }
    int out_of_stage[3];
    for (int sn = 0; sn < nstages; sn++) {
        out_of_stage[sn] = stage_output_buffer(L, sn + 1, alpha + sn * nstages);
        if (out_of_stage[sn] < 0) return set_error(-43, "a third stage needs one alpha == 0 among the older states (three buffers per patch)");
    }
    /* two compute streams (HB2_LEVEL_STREAMS=1 keeps one): the second one with its own plans, created on first use */
    static const int nstreams = [] {
        const char* v = getenv("HB2_LEVEL_STREAMS");
        return (v && *v) ? atoi(v) : 2;
    }();
    const bool two = nstreams > 1 && np > 1;
    if (two && !L->stream2) {
        HB2L_CUDA(cudaStreamCreateWithFlags(&L->stream2, cudaStreamNonBlocking));
        for (size_t k = 0; k < L->shapes.size(); k++) {
            hb2_patch_desc d = L->model;
            for (int a = 0; a < 3; a++) d.n[a] = L->shapes[k][a];
            d.device = L->device;
            hb2_plan_t plan = nullptr;
            int rc = hb2_plan_create(&d, &plan);
            if (rc) return rc;
            hb2_plan_set_stream(plan, (void*)L->stream2);
            L->plans2.push_back(plan);
        }
        L->ev_task.resize((size_t)3 * np);
        for (auto& e : L->ev_task) HB2L_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        if (slabs) HB2L_CUDA(cudaStreamWaitEvent(L->stream2, L->ev_done[0], 0));
    } else if (two && slabs) {
        HB2L_CUDA(cudaStreamWaitEvent(L->stream2, L->ev_done[0], 0));
    }
    for (const Task& t : tasks) {
        const int p = t.p, sn = t.sn;
        const double* a = alpha + sn * nstages;
        const double* b = beta + sn * nstages;
        const bool second = two && (p & 1);
        cudaStream_t st = second ? L->stream2 : L->stream;
        if (sn == 0) {
            HB2L_CUDA(cudaStreamWaitEvent(st, L->ev_up[p], 0));
            if (slabs && !edges_all && znext[p] >= 0) HB2L_CUDA(cudaStreamWaitEvent(st, L->ev_lo[znext[p]], 0));
            if (!slabs)
                for (int q : L->sources[p]) HB2L_CUDA(cudaStreamWaitEvent(st, L->ev_up[q], 0));
        } else if (two) {
            /* stage sn of patch p reads stage sn - 1 of p and of its sources; those on the other stream need an event */
            for (int q : L->sources[p])
                if (((q & 1) != 0) != second) HB2L_CUDA(cudaStreamWaitEvent(st, L->ev_task[(size_t)(sn - 1) * np + q], 0));
        }
        if (two && sn == nstages - 1 && nstages == 3) {
            /* the last stage overwrites U^(1) of p, which the second stage of p's neighbours read (ghost fill): those on the
             * same stream are ordered, those on the other stream are the sources handled above (stage sn - 1 = second stage) */
        }
        if (trace) HB2L_CUDA(cudaEventRecord(L->tr[(size_t)1 + 5 * np + (size_t)sn * np + p], st));
        const int nd = L->desc_begin[p + 1] - L->desc_begin[p];
        if (nd > 0) {
            dim3 grid(fill_blocks(L->desc_max_cells[p]), (unsigned)nd);
            k_level_fill<<<grid, 256, 0, st>>>(L->d_desc + L->desc_begin[p], L->d_ptrs[L->where[sn]], L->ncomp);
            L->launches++;
        }
        int rc = advance_stage_patch_on(L, p, sn + 1, a, b, dt, second ? L->plans2[L->patches[p].shape] : L->plans[L->patches[p].shape]);
        if (rc) return rc;
        if (two) HB2L_CUDA(cudaEventRecord(L->ev_task[(size_t)sn * np + p], st));
        if (trace) HB2L_CUDA(cudaEventRecord(L->tr[(size_t)1 + np + (size_t)sn * np + p], st));
        if (sn == nstages - 1) {
            const LevelPatch& P = L->patches[p];
            long long off, cnt;
            transfer_range(L, P, &off, &cnt);
            HB2L_CUDA(cudaEventRecord(L->ev_done[p], st));
            HB2L_CUDA(cudaStreamWaitEvent(L->copy_out, L->ev_done[p], 0));
            for (int c = 0; c < L->ncomp; c++)
                HB2L_CUDA(cudaMemcpyAsync(U_host[(size_t)p * L->ncomp + c] + off, P.S[out_of_stage[sn]] + (size_t)c * P.ncell_g + off,
                                          sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, L->copy_out));
            if (trace) HB2L_CUDA(cudaEventRecord(L->tr[(size_t)1 + 4 * np + p], L->copy_out));
        }
    }
    HB2L_CUDA(cudaGetLastError());
    HB2L_CUDA(cudaStreamSynchronize(L->copy_out));
    HB2L_CUDA(cudaStreamSynchronize(L->stream));
    if (L->stream2) HB2L_CUDA(cudaStreamSynchronize(L->stream2));
    if (trace) {
        fprintf(stderr, "[hb2_level_advance_host trace] ms since the first upload was issued: patch planes upload_done stage_done... download_done\n");
        for (int p = 0; p < np; p++) {
            float up = 0.f, dl = 0.f, tk[3] = {0.f, 0.f, 0.f}, ts[3] = {0.f, 0.f, 0.f};
            for (int sn = 0; sn < nstages; sn++) cudaEventElapsedTime(&ts[sn], L->tr[0], L->tr[(size_t)1 + 5 * np + (size_t)sn * np + p]);
            cudaEventElapsedTime(&up, L->tr[0], L->tr[1 + p]);
            for (int sn = 0; sn < nstages; sn++) cudaEventElapsedTime(&tk[sn], L->tr[0], L->tr[(size_t)1 + np + (size_t)sn * np + p]);
            cudaEventElapsedTime(&dl, L->tr[0], L->tr[(size_t)1 + 4 * np + p]);
            fprintf(stderr, "  %2d %4d  up %7.2f  s0 %7.2f-%7.2f  s1 %7.2f-%7.2f  s2 %7.2f-%7.2f  dl %7.2f\n", p, L->patches[p].n[L->dim - 1], up,
                    ts[0], tk[0], ts[1], tk[1], ts[2], tk[2], dl);
        }
    }
    for (int sn = 0; sn < nstages; sn++) {
        int rc = hb2_level_end_stage(L, sn + 1, alpha + sn * nstages, sn == nstages - 1 ? 1 : 0);
        if (rc) return rc;
    }
    return 0;
}

/* Euler::computeSpectralRadiusesAndStableDtOnPatch over the rank's patches: out_host[0..2] = max spectral radius per
 * direction, out_host[3] = max of their sum (stable dt = CFL / out_host[3]); MAX-all-reduce across ranks is the caller's. */
int hb2_level_max_wave_speed(hb2_level_t L, double out_host[4])
{
    if (!L || !out_host) return set_error(-1, "hb2_level_max_wave_speed: null argument");
    HB2L_CUDA(cudaSetDevice(L->device));
    const size_t np = L->patches.size();
    for (size_t pi = 0; pi < np; pi++) {
        const LevelPatch& P = L->patches[pi];
        const double* Q[HB2_MAX_COMP];
        for (int c = 0; c < L->ncomp; c++) Q[c] = P.S[L->where[0]] + (size_t)c * P.ncell_g;
        int rc = hb2_max_wave_speed_dev(L->plans[P.shape], Q, L->d_sr + 4 * pi);
        if (rc) return rc;
    }
    std::vector<double> h(4 * np);
    HB2L_CUDA(cudaMemcpyAsync(h.data(), L->d_sr, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, L->stream));
    HB2L_CUDA(cudaStreamSynchronize(L->stream));
    for (int k = 0; k < 4; k++) {
        out_host[k] = 0.0;
        for (size_t pi = 0; pi < np; pi++) out_host[k] = h[4 * pi + k] > out_host[k] ? h[4 * pi + k] : out_host[k];
    }
    return 0;
}

}  // extern "C"
