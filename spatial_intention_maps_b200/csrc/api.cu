// C-ABI of the simq library (include/simq.h): context, layouts, and the orchestration of the
// forward / backward / tail / optimiser kernels for networks.FCN (reference networks.py:6-26,
// resnet.py:50-120) and train.train (train.py:108-141).
#include "../../include/simq.h"
#include "kernels.h"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

thread_local long long g_simq_launches = 0;
static thread_local char g_err[512] = "";

void simq_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* simq_last_error(void) { return g_err; }
extern "C" int simq_version(void) { return 1; }

// ------------------------------------------------------------------------------------------------
// per-kernel-class timing with CUDA events on the launching stream
// ------------------------------------------------------------------------------------------------
struct ProfRec { cudaEvent_t e0, e1; int cls; double flops, issued; };
// (thread-local: simq_profile has no context argument; a profile belongs to the thread that enabled it)
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec> g_prof;
static thread_local std::vector<cudaEvent_t> g_prof_pool;
void prof_enable(bool on) { g_prof_on = on; }
void prof_mark(int cls, bool begin, double flops, cudaStream_t s, double issued) {
    if (!g_prof_on) return;
    cudaEvent_t e;
    if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    if (begin) { ProfRec r; r.e0 = e; r.e1 = nullptr; r.cls = cls; r.flops = flops; r.issued = issued; g_prof.push_back(r); }
    else if (!g_prof.empty()) g_prof.back().e1 = e;
}
static thread_local double g_prof_issued[PROF_CLASSES] = {0, 0};
int prof_collect(double* ms, double* flops, long long* launches) {
    for (int i = 0; i < PROF_CLASSES; ++i) { ms[i] = 0; flops[i] = 0; launches[i] = 0; g_prof_issued[i] = 0; }
    for (auto& r : g_prof) {
        if (!r.e1) { g_prof_pool.push_back(r.e0); continue; }
        float t = 0;
        SIMQ_CUDA(cudaEventSynchronize(r.e1));
        SIMQ_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms[r.cls] += t; flops[r.cls] += r.flops; launches[r.cls] += 1; g_prof_issued[r.cls] += r.issued;
        g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1);
    }
    g_prof.clear();
    return 0;
}
extern "C" int simq_profile_issued(double* issued) {      // of the classes collected by the last simq_profile call
    if (!issued) return 1;
    for (int i = 0; i < PROF_CLASSES; ++i) issued[i] = g_prof_issued[i];
    return 0;
}
extern "C" int simq_profile(int enable, double* ms, double* flops, long long* launches) {
    int rc = 0;
    if (ms && flops && launches) rc = prof_collect(ms, flops, launches);
    prof_enable(enable != 0);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// network description: the 70 trainable tensors / 22 BatchNorms of networks.FCN in state_dict order
// ------------------------------------------------------------------------------------------------
struct ConvP { int w; int cin, cout, k; };        // index of the weight in the param list
struct BnP { int gamma; int ch; int idx; };       // index of gamma (beta = gamma + 1), BN ordinal
struct BlockP { ConvP c1, c2, ds; BnP b1, b2, bds; bool has_ds; int cin, planes; };

struct NetDesc {
    int C, A;
    std::vector<int64_t> poff;      // 71
    std::vector<int64_t> bnoff;     // 23
    ConvP stem; BnP stem_bn;
    BlockP blk[8];
    ConvP h1, h2, h3; int h1_bias, h2_bias, h3_bias; BnP hbn1, hbn2;
};

static void build_desc(NetDesc& d, int C, int A) {
    d.C = C; d.A = A;
    d.poff.clear(); d.bnoff.clear();
    int64_t po = 0, bo = 0; int np = 0, nb = 0;
    auto add_param = [&](int64_t n) { d.poff.push_back(po); po += n; return np++; };
    auto add_conv = [&](int cin, int cout, int k) { ConvP c; c.cin = cin; c.cout = cout; c.k = k; c.w = add_param((int64_t)cout * cin * k * k); return c; };
    auto add_bn = [&](int ch) { BnP b; b.ch = ch; b.gamma = add_param(ch); add_param(ch); b.idx = nb++; d.bnoff.push_back(bo); bo += 2 * ch; return b; };
    d.stem = add_conv(C, 64, 7);
    d.stem_bn = add_bn(64);
    int inpl = 64;
    const int planes_of[4] = {64, 128, 256, 512};
    for (int li = 0; li < 4; ++li)
        for (int b = 0; b < 2; ++b) {
            BlockP& B = d.blk[li * 2 + b];
            int planes = planes_of[li];
            B.cin = b == 0 ? inpl : planes; B.planes = planes;
            B.c1 = add_conv(B.cin, planes, 3); B.b1 = add_bn(planes);
            B.c2 = add_conv(planes, planes, 3); B.b2 = add_bn(planes);
            B.has_ds = (b == 0 && inpl != planes);
            if (B.has_ds) { B.ds = add_conv(inpl, planes, 1); B.bds = add_bn(planes); }
            if (b == 1) inpl = planes;
        }
    d.h1 = add_conv(512, 128, 1); d.h1_bias = add_param(128); d.hbn1 = add_bn(128);
    d.h2 = add_conv(128, 32, 1); d.h2_bias = add_param(32); d.hbn2 = add_bn(32);
    d.h3 = add_conv(32, A, 1); d.h3_bias = add_param(A);
    d.poff.push_back(po); d.bnoff.push_back(bo);
}

extern "C" int simq_layout(int C, int A, int64_t* n_params, int64_t* n_bn, int64_t* param_offsets, int64_t* bn_offsets) {
    if (C < 1 || C > 64 || A < 1 || A > 2) { simq_set_error("simq_layout: C=%d A=%d unsupported", C, A); return 1; }
    NetDesc d; build_desc(d, C, A);
    if (n_params) *n_params = d.poff.back();
    if (n_bn) *n_bn = d.bnoff.back();
    if (param_offsets) memcpy(param_offsets, d.poff.data(), sizeof(int64_t) * d.poff.size());
    if (bn_offsets) memcpy(bn_offsets, d.bnoff.data(), sizeof(int64_t) * d.bnoff.size());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct ActSet {
    float* raw0; Split a0; Split acol;       // acol: im2col of the stem input [B*2304][Kp]
    unsigned char* a0_amax;                  // saved set only: window position of every max-pool output's first maximum (pool_bwd_kernel)
    struct { float *raw1, *raw2, *rawd; Split b1, out; } blk[8];
    float* raw_h1; Split u1; float* raw_h2; float* t;
    float* bnstat;                   // [22][4][MAX_CH] : mean, invstd, scale, shift
    int B; int training; bool valid;
};

struct PackedSet {                   // split-bf16 shadows of the conv weights of one parameter vector
    Split fwd[21], bwd[21];          // 16 3x3 + 3 downsample + head conv1 + head conv2, in network order
    Split stem;                      // resnet18.conv1 as [64][Kp]
    PackEntry* table; int n_table; long long table_total;     // device table for the one-launch packing kernel
    const float* key; uint64_t version; bool used;
};

struct simq_ctx {
    int device, maxB, backend;
    int terms;                       // 3 = parity mode (default), 1 = bf16 fast mode (simq_set_precision)
    int terms_dgrad, terms_wgrad;    // backward GEMMs: = terms, or 2 (dy contributes its hi plane only; simq_set_backward_terms)
    int dgrad2_min_planes;           // two-term dgrad only for the residual blocks with at least this many planes
    NetDesc d;
    char* pool; size_t pool_bytes, pool_used;
    ActSet set[3];                   // 0: saved (differentiated forward), 1: scratch (no-grad forwards), 2: eval-only (concurrent target pass)
    PackedSet packed[2]; int packed_next;
    // scratch
    float* partials; float* sums; double* dpartials;
    BnEntry* bn_table;               // device table of the 22 BatchNorms (one-launch eval affine)
    float *G[2], *g_mid, *du1, *dt, *dz0, *dy0, *hp, *wscratch, *stem_partials;
    Split dyA, dyB, dy2h, dy0s; float *stem_tmp, *h2_tmp;
    // second lane (see "lanes" below): its own column-sum partials and split-K / wgrad scratch, the ping-pong partner of
    // dyA, and the stash of the deferred running-statistics update of the s' pass
    float *partials2, *wscratch2, *wscratch3; Split dyA2, dyA3; double* bn_defer;
    cudaStream_t aux_stream, aux2_stream; cudaEvent_t ev_pool[32]; int ev_next; cudaEvent_t ev_done[5];
    int lanes_mode;                  // -1: read SIMQ_LANES on first use; 0 serial schedule; 1 two lanes
    cudaEvent_t next_ready;          // one-shot (simq_set_next_state_event): the next train step's s' passes wait for it
    float *q_s, *q_no, *q_nt, *dq, *per_sample; long long* best;
    int* dev_err;                    // device error word (SIMQ_DEVERR_*), read by simq_check_device_errors
    struct { int g_swapped, cur, g_parts; bool valid; } carry;      // backward state handed from phase 1 to phase 2 (simq_train_step_phase)
    unsigned int* tickets;           // per-lane "CTAs finished" counters of the conv epilogues' fused statistics reduction (zero between launches)
    long long launch_total;          // kernels launched through this context (entry points credit their launches: LaunchScope)
    // whole-step CUDA graphs (simq_train_step): one per distinct argument tuple, LRU of 8
    struct GraphEntry { std::vector<uint64_t> key; cudaGraphExec_t exec; long long launches; uint64_t last_use; };
    std::vector<GraphEntry> graphs;
    cudaStream_t side_stream; cudaEvent_t ev_in, ev_out;
    int graph_mode;                  // -1: read SIMQ_GRAPH on first use; 0 off; 1 on
    bool step_warm;                  // an eager step has run (function attributes set, tensor-map encoder resolved)
    uint64_t pack_epoch, graph_clock; // pack_epoch: bumped whenever a packed-weight slot changes owner
    int graph_misses;                // consecutive captures without a replay: a caller whose arguments change every call
                                     // (e.g. a per-step learning-rate schedule) is better served by eager launches
};

// credits the launches of one entry point (counted per thread by SIMQ_LAUNCH_CHECK) to the context it was called on
struct LaunchScope {
    simq_ctx* c; long long l0;
    explicit LaunchScope(simq_ctx* ctx) : c(ctx), l0(g_simq_launches) {}
    ~LaunchScope() { if (c) c->launch_total += g_simq_launches - l0; }
};

// Default backward operand scheme of parity mode (accepted against the float64 twin): weight-gradient GEMMs
// and the layer-4 input-gradient convolutions use dy's hi plane only (2 MMAs per product).  SIMQ_BWD_TERMS=3 restores 3 everywhere.
static void default_backward_terms(simq_ctx* c);

static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

template <typename T>
static T* carve(simq_ctx* c, size_t count, bool dry) {
    size_t bytes = align256(count * sizeof(T));
    T* p = dry ? nullptr : reinterpret_cast<T*>(c->pool + c->pool_used);
    c->pool_used += bytes;
    return p;
}
static Split carve_split(simq_ctx* c, size_t count, bool dry) {
    Split s; s.hi = carve<bf16>(c, count, dry); s.lo = carve<bf16>(c, count, dry); return s;
}

static void conv_list(const NetDesc& d, std::vector<ConvP>& out) {
    out.clear();
    for (int b = 0; b < 8; ++b) { out.push_back(d.blk[b].c1); out.push_back(d.blk[b].c2); if (d.blk[b].has_ds) out.push_back(d.blk[b].ds); }
    out.push_back(d.h1); out.push_back(d.h2);
}
static int conv_slot(const NetDesc& d, int w_index) {
    std::vector<ConvP> l; conv_list(d, l);
    for (size_t i = 0; i < l.size(); ++i) if (l[i].w == w_index) return (int)i;
    return -1;
}

static void carve_all(simq_ctx* c, bool dry) {
    const size_t B = c->maxB, R25 = B * IMG25, R48 = B * 2304;
    const NetDesc& d = c->d;
    c->pool_used = 0;
    for (int s = 0; s < 3; ++s) {
        ActSet& S = c->set[s];
        const bool eval_only = s == 2;             // folded-BN passes never write raw1 / raw2 and borrow another set's im2col
        S.raw0 = carve<float>(c, R48 * 64, dry);
        S.a0 = carve_split(c, R25 * 64, dry);
        S.a0_amax = s == 0 ? carve<unsigned char>(c, R25 * 64, dry) : nullptr;
        if (eval_only) { S.acol.hi = nullptr; S.acol.lo = nullptr; }
        else S.acol = carve_split(c, R48 * stem_kp(d.C), dry);
        for (int b = 0; b < 8; ++b) {
            size_t n = R25 * d.blk[b].planes;
            S.blk[b].raw1 = eval_only ? nullptr : carve<float>(c, n, dry);
            S.blk[b].raw2 = eval_only ? nullptr : carve<float>(c, n, dry);
            S.blk[b].rawd = d.blk[b].has_ds ? carve<float>(c, n, dry) : nullptr;
            S.blk[b].b1 = carve_split(c, n, dry);
            S.blk[b].out = carve_split(c, n, dry);
        }
        S.raw_h1 = carve<float>(c, R25 * 128, dry);
        S.u1 = carve_split(c, R48 * 128, dry);
        S.raw_h2 = carve<float>(c, R48 * 32, dry);
        S.t = carve<float>(c, R48 * 2, dry);
        S.bnstat = carve<float>(c, 22 * 4 * MAX_CH, dry);
        S.valid = false; S.B = 0; S.training = 0;
    }
    std::vector<ConvP> convs; conv_list(d, convs);
    for (int p = 0; p < 2; ++p) {
        for (size_t i = 0; i < convs.size(); ++i) {
            size_t n = (size_t)convs[i].cout * convs[i].cin * convs[i].k * convs[i].k;
            c->packed[p].fwd[i] = carve_split(c, n, dry);
            c->packed[p].bwd[i] = carve_split(c, convs[i].cout < 64 ? n / convs[i].cout * 64 : n, dry);   // dgrad K padded to 64 (head conv2)
        }
        c->packed[p].stem = carve_split(c, (size_t)64 * stem_kp(d.C), dry);
        c->packed[p].table = carve<PackEntry>(c, 32, dry);
        c->packed[p].key = nullptr; c->packed[p].version = 0; c->packed[p].used = false;
    }
    c->packed_next = 0;
    {
        size_t need = (size_t)STAT_BLOCKS * 3 * MAX_CH, conv_need = (size_t)umma_conv_m_tiles((long long)R48) * 2 * MAX_CH;
        c->partials = carve<float>(c, need > conv_need ? need : conv_need, dry);
    }
    {
        size_t need = (size_t)STAT_BLOCKS * 3 * MAX_CH, conv_need = (size_t)umma_conv_m_tiles((long long)R48) * 2 * MAX_CH;
        c->partials2 = carve<float>(c, need > conv_need ? need : conv_need, dry);
    }
    c->sums = carve<float>(c, 3 * MAX_CH, dry);
    c->dpartials = carve<double>(c, 1024, dry);
    c->bn_table = carve<BnEntry>(c, 32, dry);
    c->G[0] = carve<float>(c, R25 * 512, dry);
    c->G[1] = carve<float>(c, R25 * 512, dry);
    c->g_mid = carve<float>(c, R25 * 512, dry);
    c->du1 = carve<float>(c, R48 * 128, dry);
    c->dt = carve<float>(c, R48 * 2, dry);
    c->dz0 = carve<float>(c, R48 * 64, dry);
    c->dy0 = carve<float>(c, R48 * 64, dry);
    c->hp = carve<float>(c, (size_t)STAT_BLOCKS * 6 * 32, dry);
    c->wscratch = carve<float>(c, umma_wgrad_scratch_floats(), dry);
    c->wscratch2 = carve<float>(c, umma_wgrad_scratch_floats(), dry);
    c->wscratch3 = carve<float>(c, umma_wgrad_scratch_floats(), dry);
    c->bn_defer = carve<double>(c, (size_t)SIMQ_N_BN * 2 * MAX_CH, dry);
    c->stem_partials = carve<float>(c, stem_wgrad_partial_floats(d.C), dry);
    c->dyA = carve_split(c, R25 * 512, dry);
    c->dyA2 = carve_split(c, R25 * 512, dry);
    c->dyA3 = carve_split(c, R25 * 512, dry);
    c->dyB = carve_split(c, R25 * 512, dry);
    c->dy2h = carve_split(c, R48 * HEAD2_DY_STRIDE, dry);
    c->h2_tmp = carve<float>(c, (size_t)HEAD2_DY_STRIDE * 128, dry);
    c->dy0s = carve_split(c, R48 * 64, dry);
    c->stem_tmp = carve<float>(c, (size_t)64 * stem_kp(d.C), dry);
    c->q_s = carve<float>(c, B * 2 * 9216, dry);
    c->q_no = carve<float>(c, B * 2 * 9216, dry);
    c->q_nt = carve<float>(c, B * 2 * 9216, dry);
    c->dq = carve<float>(c, B * 2 * 9216, dry);
    c->per_sample = carve<float>(c, B * 2, dry);
    c->best = carve<long long>(c, B, dry);
    c->dev_err = carve<int>(c, 64, dry);
    c->tickets = carve<unsigned int>(c, 64, dry);
}

extern "C" int simq_ctx_create(simq_ctx** out, int device, int C, int A, int max_batch) {
    if (!out) { simq_set_error("simq_ctx_create: out is NULL"); return 1; }
    *out = nullptr;
    if (C < 1 || C > 64 || A < 1 || A > 2 || max_batch < 1) { simq_set_error("simq_ctx_create: C=%d A=%d max_batch=%d unsupported", C, A, max_batch); return 1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { simq_set_error("simq_ctx_create: no CUDA device (the simq path has no CPU fallback)"); return 2; }
    SIMQ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SIMQ_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { simq_set_error("simq_ctx_create: device sm_%d%d is not sm_100 (B200)", prop.major, prop.minor); return 2; }
    simq_ctx* c = new simq_ctx();
    c->device = device; c->maxB = max_batch; c->backend = SIMQ_BACKEND_UMMA; c->terms = 3; default_backward_terms(c);
    build_desc(c->d, C, A);
    c->pool = nullptr;
    carve_all(c, true);
    c->pool_bytes = c->pool_used;
    cudaError_t e = cudaMalloc(&c->pool, c->pool_bytes);
    if (e != cudaSuccess) { simq_set_error("simq_ctx_create: cudaMalloc(%zu) -> %s", c->pool_bytes, cudaGetErrorString(e)); delete c; return 1; }
    carve_all(c, false);
    e = cudaMemset(c->pool, 0, c->pool_bytes);
    if (e != cudaSuccess) { simq_set_error("simq_ctx_create: memset -> %s", cudaGetErrorString(e)); cudaFree(c->pool); delete c; return 1; }
    {   // device table of the BatchNorms, in BN-ordinal order
        const NetDesc& d = c->d;
        std::vector<BnEntry> host(SIMQ_N_BN);
        auto put = [&](const BnP& b, int bias_param) {
            BnEntry& E = host[b.idx];
            E.gamma_off = d.poff[b.gamma]; E.bias_off = bias_param >= 0 ? d.poff[bias_param] : -1; E.bn_off = d.bnoff[b.idx]; E.ch = b.ch; E.idx = b.idx;
        };
        put(d.stem_bn, -1);
        for (int b = 0; b < 8; ++b) { put(d.blk[b].b1, -1); put(d.blk[b].b2, -1); if (d.blk[b].has_ds) put(d.blk[b].bds, -1); }
        put(d.hbn1, d.h1_bias); put(d.hbn2, d.h2_bias);
        e = cudaMemcpy(c->bn_table, host.data(), sizeof(BnEntry) * host.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { simq_set_error("simq_ctx_create: table upload -> %s", cudaGetErrorString(e)); cudaFree(c->pool); delete c; return 1; }
    }
    {   // device tables of the one-launch weight packing
        std::vector<ConvP> convs; conv_list(c->d, convs);
        for (int p = 0; p < 2; ++p) {
            std::vector<PackEntry> host(convs.size());
            long long start = 0;
            for (size_t i = 0; i < convs.size(); ++i) {
                PackEntry& E = host[i];
                E.start = start; E.w_off = c->d.poff[convs[i].w]; E.cout = convs[i].cout; E.cin = convs[i].cin; E.kk = convs[i].k * convs[i].k; E.bk = convs[i].cout < 64 ? 64 : 0;
                E.fhi = c->packed[p].fwd[i].hi; E.flo = c->packed[p].fwd[i].lo; E.bhi = c->packed[p].bwd[i].hi; E.blo = c->packed[p].bwd[i].lo;
                start += (long long)(E.cout / 32) * (E.cin / 32);          // tiles of 32 x 32 (cout x cin); every conv of the net is a multiple
            }
            c->packed[p].n_table = (int)convs.size(); c->packed[p].table_total = start;
            e = cudaMemcpy(c->packed[p].table, host.data(), sizeof(PackEntry) * host.size(), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { simq_set_error("simq_ctx_create: table upload -> %s", cudaGetErrorString(e)); cudaFree(c->pool); delete c; return 1; }
        }
    }
    c->launch_total = 0; c->carry.valid = false;
    c->aux_stream = c->aux2_stream = nullptr; c->ev_next = 0; c->lanes_mode = -1; c->next_ready = nullptr;
    for (auto& e : c->ev_pool) e = nullptr;
    for (auto& e : c->ev_done) e = nullptr;
    c->side_stream = nullptr; c->ev_in = c->ev_out = nullptr; c->graph_mode = -1; c->step_warm = false; c->pack_epoch = 0; c->graph_clock = 0; c->graph_misses = 0;
    if (umma_init()) { cudaFree(c->pool); delete c; return 1; }
    *out = c;
    return 0;
}

extern "C" void simq_ctx_destroy(simq_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->aux2_stream) cudaStreamDestroy(c->aux2_stream);
    for (auto e : c->ev_pool) if (e) cudaEventDestroy(e);
    for (auto e : c->ev_done) if (e) cudaEventDestroy(e);
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    if (c->ev_out) cudaEventDestroy(c->ev_out);
    cudaFree(c->pool);
    delete c;
}

extern "C" int simq_set_backend(simq_ctx* c, int backend) {
    if (!c || (backend != SIMQ_BACKEND_UMMA && backend != SIMQ_BACKEND_FMA)) { simq_set_error("simq_set_backend: bad argument"); return 1; }
    c->backend = backend;
    return 0;
}
extern "C" int simq_set_precision(simq_ctx* c, int mode) {
    if (!c || (mode != SIMQ_PRECISION_PARITY && mode != SIMQ_PRECISION_BF16)) { simq_set_error("simq_set_precision: bad argument"); return 1; }
    c->terms = c->terms_dgrad = c->terms_wgrad = mode == SIMQ_PRECISION_BF16 ? 1 : 3;
    if (mode == SIMQ_PRECISION_PARITY) default_backward_terms(c);
    ++c->pack_epoch;                 // invalidates captured graphs (the key carries pack_epoch)
    return 0;
}
static void default_backward_terms(simq_ctx* c) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("SIMQ_BWD_TERMS"); forced = e ? atoi(e) : 0; }
    const int t = forced == 3 ? 3 : 2;
    c->terms_dgrad = t; c->terms_wgrad = t; c->dgrad2_min_planes = 512;
}
extern "C" int simq_set_backward_terms(simq_ctx* c, int dgrad_terms, int wgrad_terms, int dgrad2_min_planes) {
    if (!c || (dgrad_terms != 2 && dgrad_terms != 3) || (wgrad_terms != 2 && wgrad_terms != 3)) { simq_set_error("simq_set_backward_terms: terms must be 2 or 3"); return 1; }
    if (c->terms != 3) { simq_set_error("simq_set_backward_terms: only meaningful in parity mode"); return 1; }
    c->terms_dgrad = dgrad_terms; c->terms_wgrad = wgrad_terms; c->dgrad2_min_planes = dgrad2_min_planes;
    ++c->pack_epoch;
    return 0;
}
extern "C" size_t simq_workspace_bytes(const simq_ctx* c) { return c ? c->pool_bytes : 0; }
extern "C" int64_t simq_launch_count(const simq_ctx* c) { return c ? (int64_t)c->launch_total : 0; }

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
static inline float* bnstat(ActSet& S, int bn, int which) { return S.bnstat + ((size_t)bn * 4 + which) * MAX_CH; }
enum { BS_MEAN = 0, BS_INVSTD = 1, BS_SCALE = 2, BS_SHIFT = 3 };

// ------------------------------------------------------------------------------------------------
// lanes: independent pieces of a step run on two streams (forked / joined with events, so the same code
// serves eager launches and stream capture, where the fork becomes two branches of the CUDA graph):
//   forwards : lane A = online pass on s (saved set), lane B = online pass on s' then the target pass (scratch set)
//   backward : lane A = BatchNorm backward + dgrad chain, lane B = every weight-gradient GEMM
// so that the HBM-bound elementwise kernels of one lane run under the tensor-bound GEMMs of the other and the
// tails of the persistent GEMM kernels are filled.  Results are bit-identical to the serial schedule: each lane owns its
// scratch (column-sum partials, split-K / wgrad scratch), nothing is accumulated atomically, and the one ordered
// side effect two concurrent passes share -- the BatchNorm running statistics of the online net, updated by the s
// pass and then by the s' pass (train.py:114, :121) -- is applied for the s' pass after the join from a stash.
// SIMQ_LANES=0 or simq_set_schedule(ctx, SIMQ_SCHEDULE_SERIAL) selects the serial schedule (also used while
// per-kernel event profiling is on, so that a kernel's events time that kernel alone).
// ------------------------------------------------------------------------------------------------
struct Lane {
    cudaStream_t s;
    float* partials;      // per-tile column sums of the conv / dgrad epilogues, colstats / bn_bwd_reduce partials
    float* scratch;       // split-K partials of small forward convs; wgrad split partials
    double* defer;        // non-NULL: train-mode BatchNorm stashes its running-statistics update here (see above)
    unsigned int* ticket; // the lane's counter for the fused statistics reduction of its conv launches (ConvEpilogue::fin_ticket)
};
static bool fuse_stats_enabled() {      // SIMQ_FUSE_STATS=0: separate bn_finalize / reduce_partials launches (A/B experiments)
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("SIMQ_FUSE_STATS"); mode = e ? atoi(e) : 1; }
    return mode != 0;
}
static Lane main_lane(simq_ctx* c, cudaStream_t s) { return Lane{s, c->partials, c->wscratch2, nullptr, fuse_stats_enabled() ? c->tickets : nullptr}; }
static Lane side_lane(simq_ctx* c, cudaStream_t s) { return Lane{s, c->partials2, c->wscratch, nullptr, fuse_stats_enabled() ? c->tickets + 16 : nullptr}; }
// eval-mode passes only: folded BatchNorm needs no column-sum partials
static Lane eval_lane(simq_ctx* c, cudaStream_t s) { return Lane{s, nullptr, c->wscratch3, nullptr, nullptr}; }

static bool lanes_enabled(simq_ctx* c) {
    if (c->lanes_mode < 0) { const char* e = getenv("SIMQ_LANES"); c->lanes_mode = e ? (atoi(e) != 0) : 1; }
    return c->lanes_mode == 1 && !g_prof_on;
}
static bool main_priority_enabled() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("SIMQ_MAIN_PRIORITY"); mode = e ? (atoi(e) != 0) : 1; }
    return mode != 0;
}
// priority of a side lane: `steps_below_main` levels under the greatest priority, clamped to the device's range
// (SIMQ_PRIO_AUX / SIMQ_PRIO_AUX2 override the level for experiments)
static int lane_priority(const char* env, int steps_below_main) {
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return 0;
    const char* e = getenv(env);
    if (e) steps_below_main = atoi(e);
    const int p = greatest + steps_below_main;         // numerically larger = lower priority
    return p > least ? least : p;
}
static int lanes_init(simq_ctx* c) {
    if (c->aux_stream) return 0;
    // lane B / W (the online s' pass, then the weight gradients) below the main lane, lane C (target pass: BatchNorm folded, no
    // elementwise gaps of its own) below that: the lane that is left running alone at the end of the forward phase is the one
    // without gaps
    SIMQ_CUDA(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, main_priority_enabled() ? lane_priority("SIMQ_PRIO_AUX", 2) : 0));
    SIMQ_CUDA(cudaStreamCreateWithPriority(&c->aux2_stream, cudaStreamNonBlocking, main_priority_enabled() ? lane_priority("SIMQ_PRIO_AUX2", 5) : 0));
    for (auto& e : c->ev_pool) SIMQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : c->ev_done) SIMQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return 0;
}
// work enqueued on `to` after this call starts after everything enqueued on `from` so far
static int lane_order(simq_ctx* c, cudaStream_t from, cudaStream_t to, cudaEvent_t ev = nullptr) {
    if (from == to) return 0;
    if (!ev) { ev = c->ev_pool[c->ev_next]; c->ev_next = (c->ev_next + 1) % 32; }
    SIMQ_CUDA(cudaEventRecord(ev, from));
    SIMQ_CUDA(cudaStreamWaitEvent(to, ev, 0));
    return 0;
}
// the side lane of a call on stream s: a second stream, or s itself under the serial schedule
static int side_stream_for(simq_ctx* c, cudaStream_t s, cudaStream_t* out) {
    *out = s;
    if (!lanes_enabled(c)) return 0;
    if (lanes_init(c)) return 1;
    *out = c->aux_stream;
    return 0;
}

extern "C" int simq_set_schedule(simq_ctx* c, int mode) {
    if (!c || (mode != SIMQ_SCHEDULE_SERIAL && mode != SIMQ_SCHEDULE_LANES)) { simq_set_error("simq_set_schedule: bad argument"); return 1; }
    c->lanes_mode = mode == SIMQ_SCHEDULE_LANES ? 1 : 0;
    ++c->pack_epoch;                 // invalidates captured graphs (the key carries pack_epoch)
    return 0;
}

static int conv_any(simq_ctx* c, int backend, Split A, long long rows, int K, Split W, int N, int ntaps, float* out,
                    ConvEpilogue ep, const Lane& L, int terms = 0 /* 0: the context's forward precision */) {
    if (backend == SIMQ_BACKEND_UMMA && umma_conv_supported(K, N)) {
        UmmaTensor a{A, rows, K}, w{W, (long long)ntaps * N, K};
        ep.terms = terms ? terms : c->terms;
        // the lane's scratch is idle whenever one of its forward convs / dgrads runs: lend it to the split-K path of small problems
        if (out != L.scratch) { ep.splitk_scratch = L.scratch; ep.splitk_floats = umma_wgrad_scratch_floats(); }
        return k_conv_umma(a, w, N, ntaps, out, ep, L.s);
    }
    return k_conv_fma(A, rows, K, W, N, ntaps, out, ep, L.s);
}
static int wgrad_any(simq_ctx* c, int backend, Split dY, Split X, long long rows, int Cout, int Cin, int ntaps, float* dW,
                     const Lane& L) {
    if (backend == SIMQ_BACKEND_UMMA && umma_wgrad_supported(Cout, Cin)) {
        UmmaTensor y{dY, rows, Cout}, x{X, rows, Cin};
        return k_wgrad_umma(y, x, ntaps, dW, L.scratch, c->terms_wgrad, L.s);
    }
    return k_wgrad_fma(dY, X, rows, Cout, Cin, ntaps, dW, L.scratch, L.s);
}

// keep: a parameter vector whose packed slot must survive this call (the policy's, while the target's weights are looked up: a NEW
// target pointer -- a freshly built target network -- must take the other slot, not evict the policy's)
static PackedSet* get_packed(simq_ctx* c, const float* params, uint64_t version, cudaStream_t s, int* err, const float* keep = nullptr) {
    *err = 0;
    PackedSet* ps = nullptr;
    for (int i = 0; i < 2; ++i)
        if (c->packed[i].used && c->packed[i].key == params) ps = &c->packed[i];
    if (ps && version != 0 && ps->version == version) return ps;
    if (!ps) {
        int victim = c->packed_next;
        if (keep && c->packed[victim].used && c->packed[victim].key == keep) victim ^= 1;
        ps = &c->packed[victim]; c->packed_next = victim ^ 1; ++c->pack_epoch;
    }
    if (k_pack_all(params, ps->table, ps->n_table, ps->table_total, s)) { *err = 1; return nullptr; }
    if (k_pack_stem(params + c->d.poff[c->d.stem.w], c->d.C, stem_kp(c->d.C), ps->stem, s)) { *err = 1; return nullptr; }
    ps->key = params; ps->version = version; ps->used = true;
    return ps;
}

#define TRY(expr) do { if (expr) return 1; } while (0)

// BatchNorm forward bookkeeping for one BN: fills mean/invstd/scale/shift of set S.
// nparts > 0: the conv epilogue already left `nparts` partial rows in c->partials; 0: reduce `raw` here.
static int bn_prepare(simq_ctx* c, ActSet& S, const BnP& b, const float* raw, long long rows, double count, const float* params,
                      float* bn, int64_t* nbt, int bias_param, int training, int nparts, const Lane& L) {
    cudaStream_t s = L.s;
    const NetDesc& d = c->d;
    const float* gamma = params + d.poff[b.gamma];
    const float* beta = params + d.poff[b.gamma + 1];
    const float* bias = bias_param >= 0 ? params + d.poff[bias_param] : nullptr;
    float* rmean = bn + d.bnoff[b.idx];
    float* rvar = rmean + b.ch;
    if (training) {
        if (nparts == 0) { TRY(k_colstats(raw, rows, b.ch, L.partials, s)); nparts = STAT_BLOCKS; }
        TRY(k_bn_finalize_train(L.partials, nparts, b.ch, count, gamma, beta, bias, rmean, rvar, nbt ? (long long*)(nbt + b.idx) : nullptr,
                                L.defer ? L.defer + (size_t)b.idx * 2 * MAX_CH : nullptr,
                                bnstat(S, b.idx, BS_MEAN), bnstat(S, b.idx, BS_INVSTD), bnstat(S, b.idx, BS_SCALE),
                                bnstat(S, b.idx, BS_SHIFT), s));
    }
    // eval mode: run_forward has already filled scale/shift of every BN in one launch (k_bn_eval_affine_all)
    (void)beta; (void)bias; (void)rvar;
    return 0;
}

// conv -> BatchNorm statistics of its raw output.  With the tcgen05 back-end in training mode the conv
// epilogue emits the per-tile column sums itself (no extra pass over the raw output).
static int conv_bn(simq_ctx* c, ActSet& S, Split in, long long rows, int K, Split W, int N, int ntaps, float* raw, int pitch25,
                   const BnP& b, double count, const float* params, float* bn, int64_t* nbt, int bias_param, int training,
                   const Lane& L) {
    ConvEpilogue ep = conv_ep(pitch25);
    int nparts = 0;
    bool fused = false;
    if (training && c->backend == SIMQ_BACKEND_UMMA && umma_conv_supported(K, N)) {
        ep.stats = L.partials; ep.stat_rows_out = &nparts;
        if (L.ticket) {
            // the last CTA of the conv launch does the BatchNorm bookkeeping (bn_finalize_train_kernel's arithmetic) itself
            const NetDesc& d = c->d;
            fused = true;
            ep.fin_ticket = L.ticket; ep.fin_mode = 1; ep.fin_count = count;
            ep.fin_gamma = params + d.poff[b.gamma]; ep.fin_beta = params + d.poff[b.gamma + 1];
            ep.fin_bias = bias_param >= 0 ? params + d.poff[bias_param] : nullptr;
            ep.fin_rmean = bn + d.bnoff[b.idx]; ep.fin_rvar = ep.fin_rmean + b.ch;
            ep.fin_nbt = nbt ? (long long*)(nbt + b.idx) : nullptr;
            ep.fin_defer = L.defer ? L.defer + (size_t)b.idx * 2 * MAX_CH : nullptr;
            ep.fin_mean = bnstat(S, b.idx, BS_MEAN); ep.fin_invstd = bnstat(S, b.idx, BS_INVSTD);
            ep.fin_scale = bnstat(S, b.idx, BS_SCALE); ep.fin_shift = bnstat(S, b.idx, BS_SHIFT);
        }
    }
    TRY(conv_any(c, c->backend, in, rows, K, W, N, ntaps, raw, ep, L));
    if (fused) return 0;
    return bn_prepare(c, S, b, raw, rows, count, params, bn, nbt, bias_param, training, nparts, L);
}

static int run_forward(simq_ctx* c, PackedSet* pw, const float* params, float* bn, int64_t* nbt, const float* x, int B,
                       int x_layout, int training, ActSet& S, float* q, const Lane& L, const Split* acol_src = nullptr) {
    cudaStream_t s = L.s;      // acol_src: the stem im2col of the SAME input already expanded by another pass (skips the im2col launch)
    const NetDesc& d = c->d;
    const long long R25 = (long long)B * IMG25, R48 = (long long)B * 2304;
    const double cnt24 = (double)B * 576, cnt48 = (double)B * 2304;
    const int be = c->backend;
    S.valid = false;
    if (!training) TRY(k_bn_eval_affine_all(params, bn, c->bn_table, SIMQ_N_BN, S.bnstat, s));   // running-stat BN = per-channel affine
    // stem: conv 7x7/2 -> BN -> ReLU -> maxpool 3x3/2           (resnet.py:94-97)
    if (be == SIMQ_BACKEND_UMMA) {
        if (!acol_src) {
            if (!S.acol.hi) { simq_set_error("run_forward: this activation set has no im2col buffer"); return 1; }
            TRY(k_stem_im2col(x, x_layout, B, d.C, stem_kp(d.C), S.acol, s));
        }
        TRY(conv_bn(c, S, acol_src ? *acol_src : S.acol, R48, stem_kp(d.C), pw->stem, 64, 1, S.raw0, 0, d.stem_bn, cnt48, params, bn, nbt, -1, training, L));
    } else {
        TRY(k_stem_conv(x, x_layout, B, d.C, params + d.poff[d.stem.w], S.raw0, s));
        TRY(bn_prepare(c, S, d.stem_bn, S.raw0, R48, cnt48, params, bn, nbt, -1, training, 0, L));
    }
    TRY(k_stem_pool(S.raw0, B, bnstat(S, d.stem_bn.idx, BS_SCALE), bnstat(S, d.stem_bn.idx, BS_SHIFT), S.a0, training ? S.a0_amax : nullptr, s));
    // residual stages                                            (resnet.py:31-47, 99-102)
    Split in = S.a0;
    Split none{nullptr, nullptr};
    const bool fused_eval = !training && be == SIMQ_BACKEND_UMMA;
    // (eval-mode BN is a per-channel affine known before the conv runs: it is folded into the conv epilogues)
    for (int b = 0; b < 8; ++b) {
        const BlockP& P = d.blk[b];
        auto& A = S.blk[b];
        const int s1 = conv_slot(d, P.c1.w), s2 = conv_slot(d, P.c2.w);
        if (fused_eval) {
            // conv1 -> BN -> ReLU -> split activation in one kernel; conv2 -> BN -> (+identity) -> ReLU likewise
            ConvEpilogue e1 = conv_ep(1);
            e1.scale = bnstat(S, P.b1.idx, BS_SCALE); e1.shift = bnstat(S, P.b1.idx, BS_SHIFT); e1.relu = 1; e1.out_split = A.b1;
            TRY(conv_any(c, be, in, R25, P.cin, pw->fwd[s1], P.planes, 9, nullptr, e1, L));
            ConvEpilogue e2 = conv_ep(1);
            e2.scale = bnstat(S, P.b2.idx, BS_SCALE); e2.shift = bnstat(S, P.b2.idx, BS_SHIFT); e2.relu = 1; e2.out_split = A.out;
            if (P.has_ds) {
                ConvEpilogue ed = conv_ep(1);
                ed.scale = bnstat(S, P.bds.idx, BS_SCALE); ed.shift = bnstat(S, P.bds.idx, BS_SHIFT);
                TRY(conv_any(c, be, in, R25, P.cin, pw->fwd[conv_slot(d, P.ds.w)], P.planes, 1, A.rawd, ed, L));
                e2.add_prev = A.rawd;
            } else {
                e2.res = in;
            }
            TRY(conv_any(c, be, A.b1, R25, P.planes, pw->fwd[s2], P.planes, 9, nullptr, e2, L));
            in = A.out;
            continue;
        }
        TRY(conv_bn(c, S, in, R25, P.cin, pw->fwd[s1], P.planes, 9, A.raw1, 1, P.b1, cnt24, params, bn, nbt, -1, training, L));
        TRY(k_bn_apply(A.raw1, R25, P.planes, bnstat(S, P.b1.idx, BS_SCALE), bnstat(S, P.b1.idx, BS_SHIFT), 0, none, nullptr,
                       nullptr, nullptr, 1, A.b1, s));
        TRY(conv_bn(c, S, A.b1, R25, P.planes, pw->fwd[s2], P.planes, 9, A.raw2, 1, P.b2, cnt24, params, bn, nbt, -1, training, L));
        if (P.has_ds) {
            TRY(conv_bn(c, S, in, R25, P.cin, pw->fwd[conv_slot(d, P.ds.w)], P.planes, 1, A.rawd, 1, P.bds, cnt24, params, bn, nbt, -1,
                        training, L));
            TRY(k_bn_apply(A.raw2, R25, P.planes, bnstat(S, P.b2.idx, BS_SCALE), bnstat(S, P.b2.idx, BS_SHIFT), 2, none, A.rawd,
                           bnstat(S, P.bds.idx, BS_SCALE), bnstat(S, P.bds.idx, BS_SHIFT), 1, A.out, s));
        } else {
            TRY(k_bn_apply(A.raw2, R25, P.planes, bnstat(S, P.b2.idx, BS_SCALE), bnstat(S, P.b2.idx, BS_SHIFT), 1, in, nullptr,
                           nullptr, nullptr, 1, A.out, s));
        }
        in = A.out;
    }
    // head                                                       (networks.py:18-26)
    TRY(conv_bn(c, S, in, R25, 512, pw->fwd[conv_slot(d, d.h1.w)], 128, 1, S.raw_h1, 1, d.hbn1, cnt24, params, bn, nbt, d.h1_bias,
                training, L));
    TRY(k_head_up1(S.raw_h1, B, bnstat(S, d.hbn1.idx, BS_SCALE), bnstat(S, d.hbn1.idx, BS_SHIFT), S.u1, s));
    TRY(conv_bn(c, S, S.u1, R48, 128, pw->fwd[conv_slot(d, d.h2.w)], 32, 1, S.raw_h2, 0, d.hbn2, cnt48, params, bn, nbt, d.h2_bias,
                training, L));
    TRY(k_head_t(S.raw_h2, R48, bnstat(S, d.hbn2.idx, BS_SCALE), bnstat(S, d.hbn2.idx, BS_SHIFT), params + d.poff[d.h3.w], d.A,
                 S.t, s));
    if (q) TRY(k_head_up2(S.t, B, d.A, params + d.poff[d.h3_bias], q, s));
    S.B = B; S.training = training; S.valid = true;
    return 0;
}

static int check_fwd_args(simq_ctx* c, const void* params, const void* bn, const void* x, int B) {
    if (!c) { simq_set_error("ctx is NULL"); return 1; }
    if (!params || !bn || !x) { simq_set_error("NULL tensor pointer"); return 1; }
    if (B < 1 || B > c->maxB) { simq_set_error("batch %d outside [1, max_batch=%d]", B, c->maxB); return 1; }
    return 0;
}

extern "C" int simq_fcn_forward(simq_ctx* c, const float* params, float* bn, int64_t* nbt, const float* x, int B, int x_layout,
                                int training, int save_for_backward, float* q, uint64_t params_version, simq_stream stream) {
    TRY(check_fwd_args(c, params, bn, x, B));
    LaunchScope launch_scope(c);
    if (!q) { simq_set_error("simq_fcn_forward: q is NULL"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    int err;
    PackedSet* pw = get_packed(c, params, params_version, s, &err);
    if (err) return 1;
    return run_forward(c, pw, params, bn, nbt, x, B, x_layout, training, c->set[save_for_backward ? 0 : 1], q, main_lane(c, s));
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// nparts > 0: the dgrad epilogue that produced G already left `nparts` rows of (sum dz, sum dz*xhat) in L.partials
static int bn_backward(simq_ctx* c, ActSet& S, const BnP& b, const float* G, long long rows, double count, int mask_mode,
                       const bf16* mask_hi, const float* raw, const float* params, float* grads, int pitch25, Split dy,
                       float* dy_f32, const BnP* bd, const float* rawd, Split dyd, int nparts, const Lane& L, int hi_only = 0) {
    const NetDesc& d = c->d;
    cudaStream_t s = L.s;
    if (nparts < 0) {
        // c->sums already holds (sum dz, sum dz*xhat): the dgrad launch that produced G reduced its partial rows itself
    } else if (nparts > 0) {
        TRY(k_reduce_partials(L.partials, nparts, 2 * b.ch, c->sums, 1.0f, s));
    } else {
        TRY(k_bn_bwd_reduce(G, rows, b.ch, mask_mode, mask_hi, raw, bnstat(S, b.idx, BS_SCALE), bnstat(S, b.idx, BS_SHIFT),
                            bnstat(S, b.idx, BS_MEAN), bnstat(S, b.idx, BS_INVSTD), rawd, bd ? bnstat(S, bd->idx, BS_MEAN) : nullptr,
                            bd ? bnstat(S, bd->idx, BS_INVSTD) : nullptr, L.partials, s));
        TRY(k_reduce_partials(L.partials, STAT_BLOCKS, 3 * b.ch, c->sums, 1.0f, s));
    }
    TRY(k_bn_bwd_apply(G, rows, b.ch, mask_mode, mask_hi, raw, bnstat(S, b.idx, BS_SCALE), bnstat(S, b.idx, BS_SHIFT),
                       bnstat(S, b.idx, BS_MEAN), bnstat(S, b.idx, BS_INVSTD), params + d.poff[b.gamma], c->sums, count, pitch25, dy,
                       dy_f32, rawd, bd ? bnstat(S, bd->idx, BS_MEAN) : nullptr, bd ? bnstat(S, bd->idx, BS_INVSTD) : nullptr,
                       bd ? params + d.poff[bd->gamma] : nullptr, dyd, grads + d.poff[b.gamma], grads + d.poff[b.gamma + 1],
                       bd ? grads + d.poff[bd->gamma] : nullptr, bd ? grads + d.poff[bd->gamma + 1] : nullptr, hi_only, s));
    return 0;
}

static int bias_grad(simq_ctx* c, Split dy, long long rows, int ch, float* out, const Lane& L) {
    TRY(k_colsum_split(dy, rows, ch, L.partials, L.s));
    TRY(k_reduce_partials(L.partials, STAT_BLOCKS, ch, out, 1.0f, L.s));
    return 0;
}

// The backward pass.  Lane M (the caller's stream) runs the chain that carries the gradient from layer to layer --
// BatchNorm backward (-> dy of the conv below it) and the dgrad GEMMs; every weight-gradient GEMM only needs that dy and a
// saved activation, so it is handed to lane W.  The dy operand buffers are the hand-over points: a ring of three (dyA, dyA2, dyA3:
// lane M may fill the next two dy tensors while lane W still reads an earlier one), and before any dy buffer is rewritten lane M
// waits for the weight gradient that last read it (ev_done).  Lane W is joined before returning.
// phase 0: the whole backward.  phase 1: head + layer 4 (after it -- lane W joined -- every gradient from
// resnet18.layer4.0.conv1.weight to the end of the flat vector is final: 75 % of the bytes, which a data-parallel caller can
// all-reduce while phase 2 runs); phase 2: layers 3..1 + stem.  1 followed by 2 launches exactly the kernels of 0.
static int run_backward(simq_ctx* c, PackedSet* pw, const float* params, const float* x, int x_layout, const float* dq, int B,
                        float* grads, cudaStream_t s, int phase = 0) {
    const NetDesc& d = c->d;
    ActSet& S = c->set[0];
    if (!S.valid || !S.training || S.B != B) { simq_set_error("simq_fcn_backward: no matching training-mode saving forward (B=%d)", B); return 1; }
    const long long R25 = (long long)B * IMG25, R48 = (long long)B * 2304;
    const double cnt24 = (double)B * 576, cnt48 = (double)B * 2304;
    const int be = c->backend, A = d.A;
    Split none{nullptr, nullptr};
    cudaStream_t ws;
    TRY(side_stream_for(c, s, &ws));
    const Lane M = main_lane(c, s), W = side_lane(c, ws);
    const bool two = ws != s;
    enum { BUF_A0 = 0, BUF_A1 = 1, BUF_A2 = 2, BUF_B = 3, BUF_MISC = 4 };
    bool pending[5] = {false, false, false, false, false};
    // a ring of THREE: with two, lane M stalls on a weight gradient that lane W (lower priority) has not finished reading
    Split dyA[3] = {c->dyA, two ? c->dyA2 : c->dyA, two ? c->dyA3 : c->dyA};
    int cur = 1;                                         // index of the dyA buffer written last
    // lane M is about to overwrite dy buffer `buf`: wait for the weight gradient that read it
    auto acquire = [&](int buf) -> int {
        if (two && pending[buf]) { SIMQ_CUDA(cudaStreamWaitEvent(s, c->ev_done[buf], 0)); pending[buf] = false; }
        return 0;
    };
    // weight gradient of one conv on lane W, ordered after everything lane M has enqueued so far (dy is complete)
    auto wgrad_on_w = [&](int buf, Split dY, Split X, long long rows, int Cout, int Cin, int ntaps, float* dW) -> int {
        TRY(lane_order(c, s, ws));
        TRY(wgrad_any(c, be, dY, X, rows, Cout, Cin, ntaps, dW, W));
        if (two) { SIMQ_CUDA(cudaEventRecord(c->ev_done[buf], ws)); pending[buf] = true; }
        return 0;
    };
    float* G = c->G[0];
    float* Gn = c->G[1];
    int g_parts = 0;                 // partial rows already available for the BN consuming G
    ConvEpilogue ep25 = conv_ep(1);
    // With the tcgen05 back-end the dgrad launch that PRODUCES a gradient also reduces the BatchNorm-backward sums of
    // the BN that will consume it (mask = sign of the saved activation, xhat from the saved raw conv output); blocks
    // with a downsample branch need a third sum and keep the separate reduction.
    const bool fuse = be == SIMQ_BACKEND_UMMA;
    int nparts_fused = 0;            // partial rows the last BN-sum-fusing dgrad launch wrote (ConvEpilogue::stat_rows_out);
                                     // handed to bn_backward as -1 when that launch also reduced them into c->sums (fin_mode 2)
    auto parts_of = [&](const ConvEpilogue& e) { return !e.bn_raw ? 0 : e.fin_ticket ? -1 : nparts_fused; };
    // (only where the main loop is long enough -- K*taps >= 2304 -- to hide the extra epilogue loads behind the MMAs)
    auto with_bn_sums = [&](ConvEpilogue e, const BnP& bnp, const float* raw, const bf16* mask, int N, int K, int taps = 9) {
        if (fuse && umma_conv_supported(K, N) && K * taps >= 2304) {
            e.stats = M.partials; e.bn_raw = raw; e.bn_mask = mask; e.stat_rows_out = &nparts_fused;
            if (M.ticket) { e.fin_ticket = M.ticket; e.fin_mode = 2; e.fin_out = c->sums; }
            e.bn_mean = bnstat(S, bnp.idx, BS_MEAN); e.bn_invstd = bnstat(S, bnp.idx, BS_INVSTD);
        }
        return e;
    };
    const int PHASE_SPLIT_BLOCK = 5;                     // phase 1 ends after block 6 (layer4.0), phase 2 starts at block 5 (layer3.1)
    if (phase == 2) {
        if (!c->carry.valid) { simq_set_error("backward phase 2 without phase 1"); return 1; }
        if (c->carry.g_swapped) { float* t = G; G = Gn; Gn = t; }
        cur = c->carry.cur; g_parts = c->carry.g_parts;
        c->carry.valid = false;
    }
    if (phase != 2) {
    // ---- head: upsample adjoint, conv3 + BN2 + conv2 ----
    TRY(k_up2_adj(dq, B, A, c->dt, s));
    const int hb2 = d.hbn2.idx;
    TRY(k_head2_reduce(c->dt, S.raw_h2, R48, A, bnstat(S, hb2, BS_SCALE), bnstat(S, hb2, BS_SHIFT), bnstat(S, hb2, BS_MEAN),
                       bnstat(S, hb2, BS_INVSTD), params + d.poff[d.h3.w], c->hp, s));
    const int K2 = (2 + 2 * A) * 32;
    TRY(k_reduce_partials(c->hp, STAT_BLOCKS, K2, c->sums, 1.0f, s));
    TRY(k_head2_scatter(c->sums, A, grads + d.poff[d.hbn2.gamma], grads + d.poff[d.hbn2.gamma + 1], grads + d.poff[d.h3.w],
                        grads + d.poff[d.h3_bias], s));
    TRY(k_head2_apply(c->dt, S.raw_h2, R48, A, bnstat(S, hb2, BS_SCALE), bnstat(S, hb2, BS_SHIFT), bnstat(S, hb2, BS_MEAN),
                      bnstat(S, hb2, BS_INVSTD), params + d.poff[d.h3.w], c->sums, cnt48, c->dy2h, s));
    // head conv2 has 32 output channels: its dy is stored 64 wide (upper half zero), so dW comes out as [64][128] with rows
    // 32..63 zero (the first 32 rows are the gradient) and the dgrad contracts over a zero-padded K of 64
    TRY(wgrad_on_w(BUF_MISC, c->dy2h, S.u1, R48, HEAD2_DY_STRIDE, 128, 1, c->h2_tmp));
    SIMQ_CUDA(cudaMemcpyAsync(grads + d.poff[d.h2.w], c->h2_tmp, sizeof(float) * 32 * 128, cudaMemcpyDeviceToDevice, ws));
    TRY(bias_grad(c, c->dy2h, R48, HEAD2_DY_STRIDE, c->sums, M));
    SIMQ_CUDA(cudaMemcpyAsync(grads + d.poff[d.h2_bias], c->sums, sizeof(float) * 32, cudaMemcpyDeviceToDevice, s));
    ConvEpilogue ep0 = conv_ep(0);
    TRY(conv_any(c, be, c->dy2h, R48, HEAD2_DY_STRIDE, pw->bwd[conv_slot(d, d.h2.w)], 128, 1, c->du1, ep0, M, c->terms));
    // ---- head: upsample adjoint, BN1 + conv1 ----
    TRY(k_up1_adj(c->du1, B, G, s));
    cur = (cur + 1) % 3; TRY(acquire(cur));
    TRY(bn_backward(c, S, d.hbn1, G, R25, cnt24, 2, nullptr, S.raw_h1, params, grads, 1, dyA[cur], nullptr, nullptr, nullptr, none, 0, M));
    TRY(wgrad_on_w(cur, dyA[cur], S.blk[7].out, R25, 128, 512, 1, grads + d.poff[d.h1.w]));
    TRY(bias_grad(c, dyA[cur], R25, 128, grads + d.poff[d.h1_bias], M));
    {
        ConvEpilogue e = d.blk[7].has_ds ? ep25 : with_bn_sums(ep25, d.blk[7].b2, S.blk[7].raw2, S.blk[7].out.hi, 512, 128, 1);
        TRY(conv_any(c, be, dyA[cur], R25, 128, pw->bwd[conv_slot(d, d.h1.w)], 512, 1, Gn, e, M, c->terms));
        g_parts = parts_of(e);
    }
    { float* t = G; G = Gn; Gn = t; }
    }   // phase != 2
    // ---- residual stages, last to first ----
    const int b_first = phase == 2 ? PHASE_SPLIT_BLOCK : 7, b_last = phase == 1 ? PHASE_SPLIT_BLOCK + 1 : 0;
    for (int b = b_first; b >= b_last; --b) {
        const BlockP& P = d.blk[b];
        auto& Ab = S.blk[b];
        Split in = b == 0 ? S.a0 : S.blk[b - 1].out;
        const int s1 = conv_slot(d, P.c1.w), s2 = conv_slot(d, P.c2.w);
        const int dterms = P.planes >= c->dgrad2_min_planes ? c->terms_dgrad : c->terms;     // simq_set_backward_terms
        // every consumer of this block's dy tensors (two dgrads, two or three wgrads) reads the hi plane only: skip the lo planes
        const int hi_only = (be == SIMQ_BACKEND_UMMA && dterms == 2 && c->terms_wgrad == 2) ? 1 : 0;
        // out = relu(bn2(raw2) + identity): dz = G * [out > 0]
        cur = (cur + 1) % 3; TRY(acquire(cur));
        if (P.has_ds) TRY(acquire(BUF_B));
        TRY(bn_backward(c, S, P.b2, G, R25, cnt24, 1, Ab.out.hi, Ab.raw2, params, grads, 1, dyA[cur], nullptr, P.has_ds ? &P.bds : nullptr,
                        P.has_ds ? Ab.rawd : nullptr, P.has_ds ? c->dyB : none, P.has_ds ? 0 : g_parts, M, hi_only));
        TRY(wgrad_on_w(cur, dyA[cur], Ab.b1, R25, P.planes, P.planes, 9, grads + d.poff[P.c2.w]));
        if (P.has_ds) TRY(wgrad_on_w(BUF_B, c->dyB, in, R25, P.planes, P.cin, 1, grads + d.poff[P.ds.w]));
        ConvEpilogue em = with_bn_sums(ep25, P.b1, Ab.raw1, Ab.b1.hi, P.planes, P.planes);
        TRY(conv_any(c, be, dyA[cur], R25, P.planes, pw->bwd[s2], P.planes, 9, c->g_mid, em, M, dterms));
        // b1 = relu(bn1(raw1))
        cur = (cur + 1) % 3; TRY(acquire(cur));
        TRY(bn_backward(c, S, P.b1, c->g_mid, R25, cnt24, 1, Ab.b1.hi, Ab.raw1, params, grads, 1, dyA[cur], nullptr, nullptr, nullptr,
                        none, parts_of(em), M, hi_only));
        TRY(wgrad_on_w(cur, dyA[cur], in, R25, P.planes, P.cin, 9, grads + d.poff[P.c1.w]));
        // gradient w.r.t. the block input = the previous block's output (or the stem's pooled output for b == 0)
        const bool consumer_fusable = b > 0 && !d.blk[b - 1].has_ds;
        ConvEpilogue ep = ep25;
        if (!P.has_ds) { ep.add_g = G; ep.add_g_mask = Ab.out.hi; }
        g_parts = 0;
        if (consumer_fusable && !P.has_ds) ep = with_bn_sums(ep, d.blk[b - 1].b2, S.blk[b - 1].raw2, S.blk[b - 1].out.hi, P.cin, P.planes);
        TRY(conv_any(c, be, dyA[cur], R25, P.planes, pw->bwd[s1], P.cin, 9, Gn, ep, M, dterms));
        if (ep.bn_raw) g_parts = parts_of(ep);            // (the launch has just reported its partial-row count)
        if (P.has_ds) {
            ConvEpilogue epd = ep25;
            epd.add_prev = Gn;
            if (consumer_fusable) epd = with_bn_sums(epd, d.blk[b - 1].b2, S.blk[b - 1].raw2, S.blk[b - 1].out.hi, P.cin, P.planes, 1);
            TRY(conv_any(c, be, c->dyB, R25, P.planes, pw->bwd[conv_slot(d, P.ds.w)], P.cin, 1, Gn, epd, M, dterms));
            if (epd.bn_raw) g_parts = parts_of(epd);
        }
        { float* t = G; G = Gn; Gn = t; }
    }
    if (phase == 1) {
        c->carry.g_swapped = G != c->G[0]; c->carry.cur = cur; c->carry.g_parts = g_parts; c->carry.valid = true;
        TRY(lane_order(c, ws, s));                        // join: layer 4's and the head's weight gradients are final
        return 0;
    }
    // ---- stem: maxpool + ReLU + BN + conv 7x7 ----
    const int sb = d.stem_bn.idx;
    TRY(k_pool_bwd(G, S.raw0, S.a0_amax, B, bnstat(S, sb, BS_SCALE), bnstat(S, sb, BS_SHIFT), c->dz0, s));
    if (be == SIMQ_BACKEND_UMMA) {
        TRY(bn_backward(c, S, d.stem_bn, c->dz0, R48, cnt48, 0, nullptr, S.raw0, params, grads, 0, c->dy0s, nullptr, nullptr, nullptr, none, 0, M));
        TRY(wgrad_on_w(BUF_MISC, c->dy0s, S.acol, R48, 64, stem_kp(d.C), 1, c->stem_tmp));
        TRY(k_strip_stem(c->stem_tmp, d.C, stem_kp(d.C), grads + d.poff[d.stem.w], ws));
    } else {
        TRY(bn_backward(c, S, d.stem_bn, c->dz0, R48, cnt48, 0, nullptr, S.raw0, params, grads, 0, none, c->dy0, nullptr, nullptr, none, 0, M));
        TRY(k_stem_wgrad(x, x_layout, B, d.C, c->dy0, c->stem_partials, grads + d.poff[d.stem.w], s));
    }
    TRY(lane_order(c, ws, s));                            // join: every weight gradient is in `grads` for what follows on s
    return 0;
}

// the internal main-lane stream (greatest priority, see run_graphed) ordered after what `s` holds so far
static int main_stream_enter(simq_ctx* c, cudaStream_t s, bool prio, cudaStream_t* cs) {
    if (!c->side_stream) {
        int least = 0, greatest = 0;
        SIMQ_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        SIMQ_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, prio ? greatest : least));
        SIMQ_CUDA(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
        SIMQ_CUDA(cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming));
    }
    *cs = c->side_stream;
    SIMQ_CUDA(cudaEventRecord(c->ev_in, s));
    SIMQ_CUDA(cudaStreamWaitEvent(*cs, c->ev_in, 0));
    return 0;
}
// ... and `s` ordered after what the main-lane stream has done
static int main_stream_leave(simq_ctx* c, cudaStream_t s) {
    SIMQ_CUDA(cudaEventRecord(c->ev_out, c->side_stream));
    SIMQ_CUDA(cudaStreamWaitEvent(s, c->ev_out, 0));
    return 0;
}

extern "C" int simq_fcn_backward(simq_ctx* c, const float* params, const float* x, int x_layout, const float* dq, int B,
                                 float* grads, simq_stream stream) {
    if (!c || !params || !x || !dq || !grads) { simq_set_error("simq_fcn_backward: NULL argument"); return 1; }
    LaunchScope launch_scope(c);
    cudaStream_t s = (cudaStream_t)stream;
    PackedSet* pw = nullptr;
    for (int i = 0; i < 2; ++i)
        if (c->packed[i].used && c->packed[i].key == params) pw = &c->packed[i];
    if (!pw) { simq_set_error("simq_fcn_backward: parameters were not seen by a forward"); return 1; }
    if (main_priority_enabled() && lanes_enabled(c)) {        // the weight-gradient lane must not outrank this one
        cudaStream_t cs;
        TRY(main_stream_enter(c, s, true, &cs));
        TRY(run_backward(c, pw, params, x, x_layout, dq, B, grads, cs));
        return main_stream_leave(c, s);
    }
    return run_backward(c, pw, params, x, x_layout, dq, B, grads, s);
}

// ------------------------------------------------------------------------------------------------
// tail, optimiser, whole step, greedy action
// ------------------------------------------------------------------------------------------------
extern "C" int simq_dqn_tail(simq_ctx* c, const float* q_s, const float* q_next_online, const float* q_next_target,
                             const int64_t* action, const float* reward, const uint8_t* nonfinal, float gamma, int B, int Bn,
                             int double_dqn, float* out2, float* dq, simq_stream stream) {
    if (!c || !q_s || !action || !reward || !nonfinal || !out2) { simq_set_error("simq_dqn_tail: NULL argument"); return 1; }
    LaunchScope launch_scope(c);
    if (B < 1 || B > c->maxB || Bn < 0 || Bn > B) { simq_set_error("simq_dqn_tail: B=%d Bn=%d", B, Bn); return 1; }
    if (Bn > 0 && (!q_next_target || (double_dqn && !q_next_online))) { simq_set_error("simq_dqn_tail: next-state Q-maps missing"); return 1; }
    return k_dqn_tail(q_s, q_next_online, q_next_target, (const long long*)action, reward, nonfinal, gamma, B, Bn, c->d.A, double_dqn,
                      c->per_sample, c->best, out2, dq, c->dev_err, (cudaStream_t)stream);
}

extern "C" int simq_check_device_errors(simq_ctx* c, simq_stream stream) {
    if (!c) { simq_set_error("simq_check_device_errors: ctx is NULL"); return 1; }
    int flags = 0;
    SIMQ_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    SIMQ_CUDA(cudaMemcpy(&flags, c->dev_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (!flags) return 0;
    SIMQ_CUDA(cudaMemset(c->dev_err, 0, sizeof(int)));
    if (flags & SIMQ_DEVERR_ACTION_RANGE) simq_set_error("action index outside [0, %d) in the replay batch (train.py:115 would raise in gather)", c->d.A * 9216);
    else simq_set_error("device error flags 0x%x", flags);
    return 1;
}

extern "C" int simq_sgd_step(simq_ctx* c, float* params, float* grads, float* momentum, float lr, float mom, float wd,
                             float clip_norm, int first_step, float* grad_norm_out, simq_stream stream) {
    if (!c || !params || !grads || !momentum) { simq_set_error("simq_sgd_step: NULL argument"); return 1; }
    LaunchScope launch_scope(c);
    return k_sgd_step(params, grads, momentum, c->d.poff.back(), lr, mom, wd, clip_norm, first_step, c->dpartials, grad_norm_out,
                      (cudaStream_t)stream);
}

static int train_step_body(simq_ctx* c, float* params, float* bn, int64_t* nbt, const float* target_params,
                           const float* target_bn, uint64_t target_version, float* grads, float* momentum, const float* s_,
                           const float* s_next, int x_layout, const int64_t* action, const float* reward,
                           const uint8_t* nonfinal, int B, int Bn, float gamma, float lr, float mom, float wd, float clip_norm,
                           int first_step, int double_dqn, int apply_update, float* out2, cudaEvent_t next_ready, cudaStream_t s,
                           int phase = 0) {
    int err;
    if (phase == 2) {                          // the rest of the backward (layers 3..1, stem) of the step phase 1 started
        PackedSet* pw2 = nullptr;
        for (int i = 0; i < 2; ++i)
            if (c->packed[i].used && c->packed[i].key == params) pw2 = &c->packed[i];
        if (!pw2) { simq_set_error("simq_train_step_phase(2): no phase 1 on these parameters"); return 1; }
        TRY(run_backward(c, pw2, params, s_, x_layout, c->dq, B, grads, s, 2));
        if (apply_update)
            TRY(k_sgd_step(params, grads, momentum, c->d.poff.back(), lr, mom, wd, clip_norm, first_step, c->dpartials, nullptr, s));
        return 0;
    }
    PackedSet* pw = get_packed(c, params, 0, s, &err);                                     // SGD changed them last step
    if (err) return 1;
    PackedSet* pt = nullptr;
    if (Bn > 0) {
        pt = get_packed(c, target_params, target_version, s, &err, params);
        if (err) return 1;
        pw = nullptr;
        for (int i = 0; i < 2; ++i)
            if (c->packed[i].used && c->packed[i].key == params) pw = &c->packed[i];
        if (!pw) { simq_set_error("simq_train_step: packed policy weights evicted"); return 1; }
    }
    // The three forwards are independent: lane A (this stream) runs the s pass, lane B the online s' pass, lane C the target
    // pass (tcgen05 back-end + Double DQN; otherwise lane B runs both s' passes, or the only one).
    cudaStream_t bs;
    TRY(side_stream_for(c, s, &bs));
    const Lane LA = main_lane(c, s);
    Lane LB = side_lane(c, bs);
    const bool two = bs != s && Bn > 0;
    const bool three = two && double_dqn && c->backend == SIMQ_BACKEND_UMMA;
    if (two) TRY(lane_order(c, s, bs));                                                    // fork (after the weight packing)
    // train.py:114  online forward on s (train-mode BN, activations kept)
    TRY(run_forward(c, pw, params, bn, nbt, s_, B, x_layout, 1, c->set[0], c->q_s, LA));
    if (Bn > 0) {
        if (next_ready) {
            // the caller uploads s' on a copy stream while the s pass already runs: only the s' lanes wait for it.  The event
            // belongs to the caller (recorded outside any capture): inside a captured step it becomes an external event-wait node
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            SIMQ_CUDA(cudaStreamIsCapturing(LB.s, &cap));
            SIMQ_CUDA(cudaStreamWaitEvent(LB.s, next_ready, cap == cudaStreamCaptureStatusActive ? cudaEventWaitExternal : 0));
        }
        const Split* shared_acol = nullptr;
        if (double_dqn && c->backend == SIMQ_BACKEND_UMMA) {
            // both s' passes read the same stem im2col: expand it once, before lane C branches off lane B
            TRY(k_stem_im2col(s_next, x_layout, Bn, c->d.C, stem_kp(c->d.C), c->set[1].acol, LB.s));
            shared_acol = &c->set[1].acol;
        }
        Lane LC = LB;
        ActSet* target_set = &c->set[1];
        if (three) {
            LC = eval_lane(c, c->aux2_stream);
            target_set = &c->set[2];
            TRY(lane_order(c, bs, c->aux2_stream));                                        // fork C off B: after the wait and the im2col
        }
        // train.py:121  online forward on s' under no_grad, still train-mode BN (updates running stats again: after the
        // s pass's update, so a concurrent s' pass stashes its batch statistics and they are applied after the join)
        if (two) LB.defer = c->bn_defer;
        if (double_dqn) TRY(run_forward(c, pw, params, bn, nbt, s_next, Bn, x_layout, 1, c->set[1], c->q_no, LB, shared_acol));
        LB.defer = nullptr;
        if (!three) LC = LB;
        // train.py:122/124  target forward, eval-mode BN
        TRY(run_forward(c, pt, target_params, (float*)target_bn, nullptr, s_next, Bn, x_layout, 0, *target_set, c->q_nt, LC, shared_acol));
        if (three) TRY(lane_order(c, c->aux2_stream, s));                                  // join C
        if (two) {
            TRY(lane_order(c, bs, s));                                                     // join B
            if (double_dqn) TRY(k_bn_running_update_all(c->bn_table, SIMQ_N_BN, c->bn_defer, bn, (long long*)nbt, s));
        }
    }
    // train.py:115-129; dL/dQ is one-hot per sample
    float* dq = c->dq;
    TRY(k_dqn_tail(c->q_s, c->q_no, c->q_nt, (const long long*)action, reward, nonfinal, gamma, B, Bn, c->d.A, double_dqn,
                   c->per_sample, c->best, out2, dq, c->dev_err, s));
    // train.py:131-135
    TRY(run_backward(c, pw, params, s_, x_layout, dq, B, grads, s, phase));
    if (apply_update && phase == 0)
        TRY(k_sgd_step(params, grads, momentum, c->d.poff.back(), lr, mom, wd, clip_norm, first_step, c->dpartials, nullptr, s));
    return 0;
}

static inline uint64_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// Runs `body(stream)` -- a fixed sequence of this library's launches -- as a CUDA graph: captured the first time a key
// is seen, replayed afterwards (LRU of 8 graphs per context).  The first graphed call of a context runs eagerly
// (one-time function attributes), and per-kernel event profiling (simq_profile) forces the eager path.
template <typename Body, typename OnReplay>
static int run_graphed(simq_ctx* c, std::vector<uint64_t> key, cudaStream_t s, Body body, OnReplay on_replay) {
    if (c->graph_mode < 0) { const char* e = getenv("SIMQ_GRAPH"); c->graph_mode = e ? (atoi(e) != 0) : 1; }
    // Lane priorities (SIMQ_MAIN_PRIORITY=0 turns them off): the step runs on an internal stream of the greatest priority, the
    // side lanes keep lower ones (lanes_init), and the graph is instantiated with per-node priorities.  When an input-gradient
    // convolution of the critical chain and a weight-gradient GEMM are both ready, the SMs go to the convolution and the GEMM
    // fills in behind it -- under the BatchNorm backward that follows, where the tensor cores used to idle (tools/timeline.py)
    // -- instead of the two time-slicing the SMs.
    const bool prio = main_priority_enabled() && lanes_enabled(c);
    const bool eager = !c->graph_mode || g_prof_on || !c->step_warm;
    // the legacy default stream cannot be captured: run on a side stream ordered after / before it by events
    const bool side = prio || (!eager && (s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread));
    cudaStream_t cs = s;
    if (side) TRY(main_stream_enter(c, s, prio, &cs));
    if (eager) {
        int rc = body(cs);
        if (!rc) c->step_warm = true;
        if (side && !rc) TRY(main_stream_leave(c, s));
        return rc;
    }
    key.push_back((uint64_t)c->backend);
    key.push_back(c->pack_epoch);
    simq_ctx::GraphEntry* ge = nullptr;
    for (auto& g : c->graphs)
        if (g.key == key) ge = &g;
    if (ge) c->graph_misses = 0;
    else if (++c->graph_misses > 16) {                          // never replays: stop paying for capture + instantiate
        c->graph_mode = 0;
        int rc = body(cs);
        if (side && !rc) TRY(main_stream_leave(c, s));
        return rc;
    }
    if (!ge) {
        const uint64_t epoch0 = c->pack_epoch;
        const long long l0 = g_simq_launches;
        SIMQ_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
        int rc = body(cs);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(cs, &graph);
        const long long captured = g_simq_launches - l0;
        g_simq_launches = l0;                                   // nothing ran yet
        if (rc || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            if (!rc) simq_set_error("stream capture failed: %s", cudaGetErrorString(e));
            return 1;
        }
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiateWithFlags(&exec, graph, prio ? cudaGraphInstantiateFlagUseNodePriority : 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { simq_set_error("cudaGraphInstantiate: %s", cudaGetErrorString(e)); return 1; }
        if (c->graphs.size() >= 8) {                            // evict the least recently used
            size_t victim = 0;
            for (size_t i = 1; i < c->graphs.size(); ++i)
                if (c->graphs[i].last_use < c->graphs[victim].last_use) victim = i;
            cudaGraphExecDestroy(c->graphs[victim].exec);
            c->graphs.erase(c->graphs.begin() + victim);
        }
        if (c->pack_epoch != epoch0) key.back() = c->pack_epoch;   // the capture itself (re)assigned a weight slot
        c->graphs.push_back({key, exec, captured, 0});
        ge = &c->graphs.back();
    }
    ge->last_use = ++c->graph_clock;
    on_replay();                                                // host-side bookkeeping the eager body would have done
    SIMQ_CUDA(cudaGraphLaunch(ge->exec, cs));
    g_simq_launches += ge->launches;
    if (side) TRY(main_stream_leave(c, s));
    return 0;
}

static bool packed_hit(simq_ctx* c, const float* params, uint64_t version) {
    for (int i = 0; i < 2; ++i)
        if (c->packed[i].used && c->packed[i].key == params && version != 0 && c->packed[i].version == version) return true;
    return false;
}
static void packed_set_version(simq_ctx* c, const float* params, uint64_t version) {
    for (int i = 0; i < 2; ++i)
        if (c->packed[i].used && c->packed[i].key == params) c->packed[i].version = version;
}

extern "C" int simq_train_step_phase(simq_ctx* c, float* params, float* bn, int64_t* nbt, const float* target_params,
                                     const float* target_bn, uint64_t target_version, float* grads, float* momentum, const float* s_,
                                     const float* s_next, int x_layout, const int64_t* action, const float* reward,
                                     const uint8_t* nonfinal, int B, int Bn, float gamma, float lr, float mom, float wd, float clip_norm,
                                     int first_step, int double_dqn, int apply_update, float* out2, int phase, simq_stream stream) {
    TRY(check_fwd_args(c, params, bn, s_, B));
    if (phase < 0 || phase > 2) { simq_set_error("simq_train_step_phase: phase %d", phase); return 1; }
    LaunchScope launch_scope(c);
    if (!target_params || !target_bn || !grads || !momentum || !action || !reward || !nonfinal || !out2) { simq_set_error("simq_train_step: NULL argument"); return 1; }
    if (Bn < 0 || Bn > B || (Bn > 0 && !s_next)) { simq_set_error("simq_train_step: Bn=%d", Bn); return 1; }
    // whether the packed target weights are reused decides whether their pack kernels are part of the graph
    const bool target_hit = packed_hit(c, target_params, target_version);
    std::vector<uint64_t> key = {1, (uint64_t)params, (uint64_t)bn, (uint64_t)nbt, (uint64_t)target_params, (uint64_t)target_bn,
                                 (uint64_t)target_hit, (uint64_t)grads, (uint64_t)momentum, (uint64_t)s_, (uint64_t)s_next, (uint64_t)x_layout,
                                 (uint64_t)action, (uint64_t)reward, (uint64_t)nonfinal, (uint64_t)B, (uint64_t)Bn, fbits(gamma), fbits(lr),
                                 fbits(mom), fbits(wd), fbits(clip_norm), (uint64_t)first_step, (uint64_t)double_dqn, (uint64_t)apply_update,
                                 (uint64_t)out2, (uint64_t)(phase == 2 ? nullptr : c->next_ready), (uint64_t)phase};
    cudaEvent_t next_ready = phase == 2 ? nullptr : c->next_ready;
    if (phase != 2) c->next_ready = nullptr;                    // one-shot
    return run_graphed(c, key, (cudaStream_t)stream,
        [&](cudaStream_t st) {
            return train_step_body(c, params, bn, nbt, target_params, target_bn, target_version, grads, momentum, s_, s_next, x_layout,
                                   action, reward, nonfinal, B, Bn, gamma, lr, mom, wd, clip_norm, first_step, double_dqn, apply_update,
                                   out2, next_ready, st, phase);
        },
        [&]() {
            if (phase != 2 && Bn > 0) packed_set_version(c, target_params, target_version);
            c->set[0].valid = true; c->set[0].B = B; c->set[0].training = 1;
            if (phase == 1) c->carry.valid = true;              // (the captured body filled the carry when the graph was built)
            if (phase == 2) c->carry.valid = false;
        });
}

extern "C" int simq_train_step(simq_ctx* c, float* params, float* bn, int64_t* nbt, const float* target_params,
                               const float* target_bn, uint64_t target_version, float* grads, float* momentum, const float* s_,
                               const float* s_next, int x_layout, const int64_t* action, const float* reward,
                               const uint8_t* nonfinal, int B, int Bn, float gamma, float lr, float mom, float wd, float clip_norm,
                               int first_step, int double_dqn, int apply_update, float* out2, simq_stream stream) {
    return simq_train_step_phase(c, params, bn, nbt, target_params, target_bn, target_version, grads, momentum, s_, s_next, x_layout, action,
                                 reward, nonfinal, B, Bn, gamma, lr, mom, wd, clip_norm, first_step, double_dqn, apply_update, out2, 0, stream);
}

extern "C" int simq_set_next_state_event(simq_ctx* c, void* event) {
    if (!c) { simq_set_error("simq_set_next_state_event: ctx is NULL"); return 1; }
    c->next_ready = (cudaEvent_t)event;
    return 0;
}

extern "C" int simq_gather_rows(const float* src, const int64_t* idx, int n, int64_t row_floats, float* dst, simq_stream stream) {
    if (n < 0 || (n > 0 && (!src || !idx || !dst)) || row_floats < 4) { simq_set_error("simq_gather_rows: bad argument"); return 1; }
    return k_gather_rows(src, (const long long*)idx, n, row_floats, dst, (cudaStream_t)stream);
}

extern "C" int simq_bce_tail(simq_ctx* c, const float* q, const float* target, int64_t target_stride, int64_t n, float* out1, float* dq,
                             simq_stream stream) {
    if (!c || !q || !target || !out1 || !dq || n < 1 || target_stride < 1) { simq_set_error("simq_bce_tail: bad argument"); return 1; }
    LaunchScope launch_scope(c);
    return k_bce_tail(q, target, target_stride, n, out1, dq, c->dpartials, (cudaStream_t)stream);
}

extern "C" int simq_intention_step(simq_ctx* c, float* params, float* bn, int64_t* nbt, float* grads, float* momentum,
                                   const float* state, int B, float lr, float mom, float wd, float clip_norm, int first_step,
                                   int apply_update, float* out1, simq_stream stream) {
    TRY(check_fwd_args(c, params, bn, state, B));
    LaunchScope launch_scope(c);
    if (!grads || !momentum || !out1) { simq_set_error("simq_intention_step: NULL argument"); return 1; }
    if (c->d.A != 1) { simq_set_error("simq_intention_step: the intention net has one output channel (A=%d)", c->d.A); return 1; }
    std::vector<uint64_t> key = {2, (uint64_t)params, (uint64_t)bn, (uint64_t)nbt, (uint64_t)grads, (uint64_t)momentum, (uint64_t)state,
                                 (uint64_t)B, fbits(lr), fbits(mom), fbits(wd), fbits(clip_norm), (uint64_t)first_step,
                                 (uint64_t)apply_update, (uint64_t)out1};
    return run_graphed(c, key, (cudaStream_t)stream,
        [&](cudaStream_t s) {
            int err;
            PackedSet* pw = get_packed(c, params, 0, s, &err);
            if (err) return 1;
            // train.py:145-148: inputs = all channels but the last, target = the last channel of the same state
            TRY(run_forward(c, pw, params, bn, nbt, state, B, SIMQ_X_NHWC_PLUS1, 1, c->set[0], c->q_s, main_lane(c, s)));
            TRY(k_bce_tail(c->q_s, state + c->d.C, c->d.C + 1, (long long)B * 9216, out1, c->dq, c->dpartials, s));     // :149-150
            TRY(run_backward(c, pw, params, state, SIMQ_X_NHWC_PLUS1, c->dq, B, grads, s));                              // :151-152
            if (apply_update)                                                                                           // :153
                TRY(k_sgd_step(params, grads, momentum, c->d.poff.back(), lr, mom, wd, clip_norm, first_step, c->dpartials, nullptr, s));
            return 0;
        },
        [&]() { c->set[0].valid = true; c->set[0].B = B; c->set[0].training = 1; });
}

extern "C" int simq_greedy_action(simq_ctx* c, const float* params, const float* bn, const float* x, int B, int x_layout,
                                  int64_t* action_out, float* q, uint64_t params_version, simq_stream stream) {
    TRY(check_fwd_args(c, params, bn, x, B));
    LaunchScope launch_scope(c);
    if (!action_out) { simq_set_error("simq_greedy_action: action_out is NULL"); return 1; }
    // the env loop calls this once per step with the same buffers: replay one graph (policies.py:47-74)
    const bool hit = packed_hit(c, params, params_version);
    std::vector<uint64_t> key = {3, (uint64_t)params, (uint64_t)bn, (uint64_t)x, (uint64_t)B, (uint64_t)x_layout, (uint64_t)action_out,
                                 (uint64_t)q, (uint64_t)hit};
    return run_graphed(c, key, (cudaStream_t)stream,
        [&](cudaStream_t s) {
            int err;
            PackedSet* pw = get_packed(c, params, params_version, s, &err);
            if (err) return 1;
            float* qq = q ? q : c->q_nt;
            TRY(run_forward(c, pw, params, (float*)bn, nullptr, x, B, x_layout, 0, c->set[1], qq, main_lane(c, s)));
            return k_argmax_rows(qq, B, (long long)c->d.A * 9216, (long long*)action_out, s);
        },
        [&]() { packed_set_version(c, params, params_version); });
}

// ------------------------------------------------------------------------------------------------
// test hooks
// ------------------------------------------------------------------------------------------------
static const char* kDebugNames[] = {"raw0", "a0"};
extern "C" const char* simq_debug_tensor_name(int id) {
    static thread_local char buf[64];
    if (id < 0) return nullptr;
    if (id < 2) return kDebugNames[id];
    if (id < 34) { const char* n[4] = {"raw1", "b1", "raw2", "out"}; snprintf(buf, sizeof(buf), "blk%d.%s", (id - 2) / 4, n[(id - 2) % 4]); return buf; }
    if (id == 34) return "raw_h1";
    if (id == 35) return "u1";
    if (id == 36) return "raw_h2";
    if (id == 37) return "t";
    if (id < 46) { snprintf(buf, sizeof(buf), "blk%d.rawd", id - 38); return buf; }
    return nullptr;
}

extern "C" int simq_debug_get(simq_ctx* c, int set, int id, int B, float* out, int64_t* chw, simq_stream stream) {
    if (!c || set < 0 || set > 1 || !out) { simq_set_error("simq_debug_get: bad argument"); return 1; }
    LaunchScope launch_scope(c);
    ActSet& S = c->set[set];
    cudaStream_t s = (cudaStream_t)stream;
    Split none{nullptr, nullptr};
    const NetDesc& d = c->d;
    int C = 0, HW = 24; const float* raw = nullptr; Split sp = none; bool p25 = true;
    if (id == 0) { raw = S.raw0; C = 64; HW = 48; p25 = false; }
    else if (id == 1) { sp = S.a0; C = 64; }
    else if (id < 34) {
        int b = (id - 2) / 4, w = (id - 2) % 4; C = d.blk[b].planes;
        if (w == 0) raw = S.blk[b].raw1; else if (w == 1) sp = S.blk[b].b1; else if (w == 2) raw = S.blk[b].raw2; else sp = S.blk[b].out;
    }
    else if (id == 34) { raw = S.raw_h1; C = 128; }
    else if (id == 35) { sp = S.u1; C = 128; HW = 48; p25 = false; }
    else if (id == 36) { raw = S.raw_h2; C = 32; HW = 48; p25 = false; }
    else if (id == 37) { raw = S.t; C = d.A; HW = 48; p25 = false; }
    else if (id < 46) { int b = id - 38; if (!d.blk[b].has_ds) { simq_set_error("simq_debug_get: block %d has no downsample", b); return 1; } raw = S.blk[b].rawd; C = d.blk[b].planes; }
    else { simq_set_error("simq_debug_get: unknown id %d", id); return 1; }
    if (chw) *chw = (int64_t)C * HW * HW;
    if (p25) return k_export_p25(raw, sp, B, C, out, s);
    return k_export_dense(raw, sp, B, C, HW, out, s);
}

extern "C" int simq_test_conv(simq_ctx* c, int backend, int mode, int B, int Cin, int Cout, int k, const float* a, const float* a2,
                              const float* w, float* out, simq_stream stream) {
    if (!c || !a || !out || (mode != 2 && !w) || (mode == 2 && !a2)) { simq_set_error("simq_test_conv: NULL argument"); return 1; }
    LaunchScope launch_scope(c);
    if (B < 1 || B > c->maxB || Cin > 512 || Cout > 512 || Cin % 16 || Cout % 16 || (k != 1 && k != 3)) { simq_set_error("simq_test_conv: bad shape"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    const long long R25 = (long long)B * IMG25;
    const int taps = k * k;
    // borrow backward scratch: dyA / dyB as the operand planes, G[0] as the p25 output, packed slot 1 for the weights
    Split X = c->dyA, Y = c->dyB;
    Split wf = c->packed[1].fwd[15], wb = c->packed[1].bwd[15];       // layer4.1.conv2 slot: 512*512*9 elements
    c->packed[1].used = false;
    Split none{nullptr, nullptr};
    ConvEpilogue ep25 = conv_ep(1);
    if (mode == 0) {
        TRY(k_import_p25(a, B, Cin, X, nullptr, s));
        TRY(k_pack_weights(w, Cout, Cin, taps, wf, wb, s));
        TRY(conv_any(c, backend, X, R25, Cin, wf, Cout, taps, c->G[0], ep25, main_lane(c, s)));
        return k_export_p25(c->G[0], none, B, Cout, out, s);
    } else if (mode == 1) {
        TRY(k_import_p25(a, B, Cout, Y, nullptr, s));
        TRY(k_pack_weights(w, Cout, Cin, taps, wf, wb, s));
        TRY(conv_any(c, backend, Y, R25, Cout, wb, Cin, taps, c->G[0], ep25, main_lane(c, s), c->terms_dgrad));
        return k_export_p25(c->G[0], none, B, Cin, out, s);
    } else {
        TRY(k_import_p25(a, B, Cin, X, nullptr, s));
        TRY(k_import_p25(a2, B, Cout, Y, nullptr, s));
        return wgrad_any(c, backend, Y, X, R25, Cout, Cin, taps, out, side_lane(c, s));
    }
}
