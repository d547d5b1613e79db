/* engine.cu -- libx265cu.so: the C ABI of include/x265cu.h on one B200.
 *
 * Host side of the engine: owns the frame slots in HBM, turns job batches into kernel launches and
 * copies results back.  All arithmetic lives in la_kernels.cuh.  There is no CPU implementation of
 * any job here: without a usable CUDA device x265cu_create fails.
 *
 * Streams: `copyStream` carries the picture uploads; `preStream` the pre-lookahead kernels (K1-K3); `stream`
 * (main) weightp scores, cuTree and every result copy, so the decisions never queue behind an upload; and
 * LA_NUM_LANES worker streams carry the search / cost batches, round-robin, so the wavefront searches of
 * consecutive batches overlap.  Ordering between them is by events: batch after the pre-lookahead stream at
 * batch_begin; main-stream work after the pre-lookahead of the slots it reads; cost jobs after the batches whose MV stores they read; main-stream readers after the batch
 * that writes the store they read; an upload after every batch that touched the slot's previous tenant.
 *
 * HBM layout of one frame slot (one cudaMalloc, 256-byte aligned sections):
 *   srcY/U/V        packed full-res picture (only needed until K1/K2 ran)
 *   planes          4 half-pel planes with margins, contiguous, exactly Lowres::buffer[0..3]
 *   intraCost, intraMode, invQscale, qpAq, qpCuTree, propagate (int32 accumulators), energy
 *   lowresCosts00, rowSatds00, stats
 *   mvStores        3*nb x { int packedMv[ncu], int mvCost[ncu] }
 *   costStores      2*nb*nb x { u16 lowresCosts[ncu], int rowSatds[bh], CostResultDev }
 */
#include "x265cu.h"
#include "la_kernels.cuh"
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <map>
#include <algorithm>
#include <new>

using namespace la;

static_assert(sizeof(HistStatsDev) == sizeof(x265cu_hist_stats), "HistStatsDev mirrors x265cu_hist_stats");
static_assert(sizeof(FrameStatsDev) == sizeof(x265cu_frame_stats), "FrameStatsDev mirrors x265cu_frame_stats");

#include <chrono>
#include <mutex>
/* host-side stopwatch for tuning (X265CU_HOST_TIMING=1 prints the totals when a context is destroyed) */
enum { HT_UPLOAD, HT_BATCH_BEGIN, HT_SEARCH_ENQ, HT_COST_ENQ, HT_BATCH_END, HT_GATHER, HT_MIRROR_RINGWAIT, HT_MIRROR_MALLOC, HT_MIRROR_REST,
       HT_STATS_WAIT, HT_RECALC_GET, HT_CUTREE, HT_WEIGHT, HT_COUNT };
static const char* const g_htNames[HT_COUNT] = { "upload", "batch_begin", "search_enqueue", "cost_enqueue", "batch_end", "gather", "mirror_ringwait",
                                                 "mirror_malloc", "mirror_rest", "stats_wait", "recalc_get", "cutree", "weight" };
static double g_ht[HT_COUNT];
static unsigned long long g_htN[HT_COUNT];
struct HostTimer
{
    int k; std::chrono::steady_clock::time_point t0;
    explicit HostTimer(int kind) : k(kind), t0(std::chrono::steady_clock::now()) {}
    ~HostTimer() { g_ht[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); g_htN[k]++; }
};

namespace {

struct SlotLayout
{
    size_t srcY, srcU, srcV, planes, intraCost, intraMode, invQ, invQ8, qpAq, qpCuTree, propagate, energy, edge, aqSums,
           lowresCosts00, rowSatds00, stats, mvStores, costStores, planes4, mvStores4, total;
    size_t mvStoreStride, costStoreStride, costRowOff, costResOff, mvStore4Stride;
};

inline size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

} // namespace

enum { LA_NUM_LANES = 16, LA_NUM_BATCHES = 48, LA_MAX_RANKS = 8, LA_CT_RING = 2048, LA_NUM_PRE = 3 };

/* One asynchronous batch of search / cost jobs (x265cu_batch_begin ... x265cu_batch_end).  Everything a batch's
 * kernels read besides the frame slots is private to it, so batches never wait for each other's buffers. */
struct Batch
{
    long long id;                       /* -1 = never used; the object serves ids id, id + LA_NUM_BATCHES, ... */
    cudaStream_t stream;                /* lane id % LA_NUM_LANES */
    cudaEvent_t begun[LA_NUM_PRE], searchDone, done;
    bool open;
    char* h_stage; char* d_stage;       /* pinned host / device copy of the job arrays */
    size_t stageCap, stageUsed;
    int* d_sync; size_t syncCap, syncUsed;      /* ticket counter + per-(job, strip) progress of every search call */
    std::vector<char*> weightScratch;   /* 4 weighted planes each */
    std::vector<void*> retiredDev, retiredHost; /* outgrown buffers, freed when the object is reused */
    std::vector<long long> waited;      /* batches whose searches this one already waits for */
    std::vector<long long> waitedDone;  /* batches this one waits for entirely */
    /* sharded stream: stores written in this batch (by their owners), exchanged at batch_end */
    struct Seg { char* ptr; size_t bytes; int root; };
    std::vector<Seg> segs;
    char* xbuf[LA_MAX_RANKS]; size_t xcap[LA_MAX_RANKS];
};

struct x265cu_ctx
{
    x265cu_config cfg;
    Geom g;
    Geom g4;                        /* --hme: geometry of the 1/16-resolution planes and their block grid (level 0) */
    x265cu_geometry geom;
    int bpp;
    SlotLayout lay;
    cudaStream_t stream;            /* main: weightp scores, cuTree, every D2H */
    cudaStream_t copyStream;        /* picture uploads, so they overlap the kernels of earlier frames */
    cudaStream_t preStreams[LA_NUM_PRE];    /* pre-lookahead kernels K1-K3: consecutive frames alternate between them, so the one-warp
                                               sequential AQ mean of a frame overlaps the streaming kernels of the next ones */
    unsigned preSeq;
    cudaEvent_t mainMark;
    /* host mirrors of decided frames (x265cu_mirror_enqueue): their own high-priority stream, a ring of requests */
    cudaStream_t mirrorStream;
    cudaStream_t gatherStream;      /* the small synchronous gathers of the decision path (cost sums, skip flags): ordered only after the
                                       batches that produced them, NOT after the cuTree backlog of earlier decisions on the main stream */
    cudaEvent_t mirrorMark;
    struct MirrorEntry { cudaEvent_t done; char* scratch; size_t cap; long long ticket; } mirror[X265CU_MIRROR_RING];
    long long nextMirror;
    std::map<void*, void*> mappedCache;         /* host pointer -> device address of page-locked memory (NULL: not page-locked) */
    std::vector<cudaEvent_t> slotMirrored;      /* per slot: the last mirror request that read it */
    std::vector<char> slotMirrorTouched;
    /* cost recalculation ahead of time (x265cu_cost_recalc_enqueue): per slot {score, rows[bh]} in device scratch + event */
    std::vector<cudaEvent_t> slotRecalc;
    std::vector<int> slotRecalcStore;           /* cost store the pending request is for, -1 = none */
    char* d_recalc; char* h_recalc;             /* [slot] x recalcStride: device scratch / mapped host copy */
    size_t recalcStride;
    std::vector<char> slotMainTouched;          /* main-stream work read the slot's current tenant */
    std::vector<cudaEvent_t> slotMainWrote;     /* per slot: the last main-stream kernel that wrote arrays a mirror reads (cuTreeFinish,
                                                   the synchronous cost recalculation) */
    std::vector<char> slotMainWroteSet;
    cudaStream_t lanes[LA_NUM_LANES];
    Batch batches[LA_NUM_BATCHES];
    long long nextBatch;
    Batch* cur;                     /* the open batch or NULL */
    std::vector<char*> slots;
    std::vector<cudaEvent_t> slotCopied, slotConsumed;   /* per slot: upload done / pre-lookahead done */
    std::vector<std::vector<long long> > slotUsers;      /* batches that read or write the slot's current tenant */
    std::vector<long long> mvWriter, costWriter;         /* [slot * n_stores + store] -> batch that writes it, -1 */
    unsigned short* d_mvcost;       /* whole table; centre at +mvcost_half */
    unsigned long long* d_executed; /* [0] search jobs, [1] cost jobs that passed their condition */
    char* d_results; size_t resultsCap;
    HistAcc* d_histAcc; HistStatsDev* h_hist; HistStatsDev* d_hist;   /* --hist-scenecut: per-slot accumulators, per-slot results in mapped host memory */
    FrameStatsDev* h_slotStats; FrameStatsDev* d_slotStats;   /* mapped host memory: every slot's statistics, written by K3's epilogue */
    char* h_mapped; char* d_mapped; size_t mappedCap;         /* mapped host memory for the small gathers */
    CutreeJobDev* h_ctJobs; CutreeJobDev* d_ctJobs;           /* mapped ring of batched cuTree propagate jobs */
    int ctRingPos, ctPending;                                 /* next free ring entry; jobs gathered but not launched */
    std::vector<char*> mainScratch;           /* weighted plane for x265cu_weight_cost_batch (main stream) */
    x265cu_counters counters;
    uint64_t searchEnq, costEnq;    /* jobs enqueued (conditional ones included) */
    int rank, nranks;               /* sharded stream (x265cu_shard_config); nranks 1 = not sharded */
    x265cu_exchange_fn exchange; void* exchangeUser;
    std::vector<int> slotOwner;
    std::vector<std::pair<char*, size_t> > xpool;    /* free exchange buffers */
    int searchWorkers;              /* worker warps per search job; 0 = default (env X265CU_SEARCH_WORKERS, for tuning) */
    int numSMs;
    int searchSmem;                 /* dynamic shared memory per search CTA (env X265CU_SEARCH_SMEM, bytes): residency cap */
    int searchLanes;                /* lanes per 8x8 block in the search kernel: 4 (default; two rows per lane, strips of 8 block rows) or
                                       8 (one row per lane, strips of 4; env X265CU_SEARCH_LANES, kept for A/B measurements) */
    int searchOneShot;              /* one ticket per search CTA instead of persistent workers (env X265CU_SEARCH_ONESHOT) */
    int costLanes;                  /* lanes per block in the grouped cost kernel: 4 (default) or 8 (env X265CU_COST_LANES, for A/B) */
    long long bigSearch[2];         /* the two most recent batches with a large search launch (see searchBatchT), -1 = none */
    bool profile;
    double profMs[X265CU_K_COUNT], profBusy[X265CU_K_COUNT];
    uint64_t profN[X265CU_K_COUNT];
    struct EvPair { cudaEvent_t a, b; int kind; };
    std::vector<EvPair> evPool;     /* non-blocking per-launch timing: resolved in x265cu_profile_get */
    size_t evUsed;
    cudaEvent_t profBase;           /* origin of the busy-interval timestamps */
    cudaEvent_t tm0, tm1;           /* x265cu_timer_* */
    /* SM partitioning (CUDA green contexts): the latency-critical short kernels the host waits for (cuTree, cost
     * recalculation, weightp scores, the mirror's unpack / de-tile) get a small SM partition of their own, the search / cost
     * lanes the rest.  Without it a 256-thread cuTree CTA queues for milliseconds until enough of the long-lived one-warp search
     * CTAs of ONE SM have retired (28 of them hold every register of an SM).  NULL = not in use (API missing, or X265CU_GREEN=0). */
    void* greenSmall; void* greenLarge;
    int greenSmallSMs, greenLargeSMs;
    char err[256];
};

namespace {

bool cudaOk(x265cu_ctx* c, cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    snprintf(c->err, sizeof(c->err), "%s: %s", what, cudaGetErrorString(e));
    return false;
}
#define CK(call) do { if (!cudaOk(c, (call), #call)) return X265CU_ERR_CUDA; } while (0)

/* Every entry point runs with the context's GPU current and puts the caller's back afterwards (a process may drive
 * several GPUs, and the caller -- torch, an encoder with its own CUDA code -- has its own idea of the current device) */
struct DeviceScope
{
    int prev, want;
    explicit DeviceScope(const x265cu_ctx* c);
    ~DeviceScope() { if (prev != want && prev >= 0) cudaSetDevice(prev); }
};

static std::mutex g_evMutex;
static std::map<int, std::vector<std::pair<cudaEvent_t, cudaEvent_t> > > g_evFree;     /* per device: timing event pairs not in use */

/* Brackets the launches of one kernel family with a pair of CUDA events on the launching stream.
 * Nothing blocks here; the pairs are resolved when the caller asks for the totals. */
struct Prof
{
    x265cu_ctx* c; int k; int idx; cudaStream_t st;
    Prof(x265cu_ctx* ctx, int kind, int launches, cudaStream_t stream = 0) : c(ctx), k(kind), idx(-1), st(stream ? stream : ctx->stream)
    {
        c->counters.kernel_launches += launches;
        c->profN[k] += launches;
        if (c->profile)
        {
            if (c->evUsed == c->evPool.size())
            {
                /* timing events are recycled across contexts of the same device (a bench step opens a new context and
                 * brackets ~2000 launches: 4000 cudaEventCreate calls per step otherwise) */
                x265cu_ctx::EvPair e;
                e.a = e.b = NULL; e.kind = 0;
                {
                    std::lock_guard<std::mutex> lock(g_evMutex);
                    std::vector<std::pair<cudaEvent_t, cudaEvent_t> >& fr = g_evFree[c->cfg.device];
                    if (!fr.empty()) { e.a = fr.back().first; e.b = fr.back().second; fr.pop_back(); }
                }
                if (!e.a) { cudaEventCreate(&e.a); cudaEventCreate(&e.b); }
                c->evPool.push_back(e);
            }
            idx = (int)c->evUsed++;
            c->evPool[idx].kind = k;
            cudaEventRecord(c->evPool[idx].a, st);
        }
    }
    ~Prof()
    {
        if (idx >= 0) cudaEventRecord(c->evPool[idx].b, st);
    }
};

void syncAll(x265cu_ctx* c)
{
    cudaStreamSynchronize(c->copyStream);
    for (int i = 0; i < LA_NUM_PRE; i++) cudaStreamSynchronize(c->preStreams[i]);
    cudaStreamSynchronize(c->stream);
    if (c->mirrorStream) cudaStreamSynchronize(c->mirrorStream);
    if (c->gatherStream) cudaStreamSynchronize(c->gatherStream);
    for (int i = 0; i < LA_NUM_LANES; i++) cudaStreamSynchronize(c->lanes[i]);
}

/* per family: sum of the launch durations, and the length of the union of the launch intervals (batches overlap) */
void resolveProfile(x265cu_ctx* c)
{
    if (!c->evUsed) return;
    syncAll(c);
    std::vector<std::pair<float, float> > iv[X265CU_K_COUNT];
    for (size_t i = 0; i < c->evUsed; i++)
    {
        float t0 = 0, t1 = 0;
        if (cudaEventElapsedTime(&t0, c->profBase, c->evPool[i].a) == cudaSuccess &&
            cudaEventElapsedTime(&t1, c->profBase, c->evPool[i].b) == cudaSuccess)
        {
            c->profMs[c->evPool[i].kind] += t1 - t0;
            iv[c->evPool[i].kind].push_back(std::make_pair(t0, t1));
        }
    }
    if (const char* path = getenv("X265CU_TIMELINE"))
    {
        /* tuning aid: every profiled launch bracket as "family,start_ms,end_ms" (ms since the context was created) */
        if (FILE* f = fopen(path, "a"))
        {
            for (int k = 0; k < X265CU_K_COUNT; k++)
                for (size_t i = 0; i < iv[k].size(); i++) fprintf(f, "%d,%.4f,%.4f\n", k, iv[k][i].first, iv[k][i].second);
            fprintf(f, "-1,0,0\n");
            fclose(f);
        }
    }
    for (int k = 0; k < X265CU_K_COUNT; k++)
    {
        std::sort(iv[k].begin(), iv[k].end());
        float end = -1e30f;
        for (size_t i = 0; i < iv[k].size(); i++)
        {
            if (iv[k][i].first > end) { c->profBusy[k] += iv[k][i].second - iv[k][i].first; end = iv[k][i].second; }
            else if (iv[k][i].second > end) { c->profBusy[k] += iv[k][i].second - end; end = iv[k][i].second; }
        }
    }
    c->evUsed = 0;
}

/* launch the cuTree propagate jobs gathered so far (consecutive unreferenced B frames) as one kernel.  Every entry
 * point that touches the main stream or recycles a slot calls this first, so gathering never reorders anything. */
void flushCutree(x265cu_ctx* c)
{
    if (!c || !c->ctPending) return;
    const int n = c->ctPending;
    c->ctPending = 0;
    c->counters.kernel_launches++; c->profN[X265CU_K_CUTREE]++;
    cutree_propagate_batch_kernel<<<dim3((c->g.ncu + 255) / 256, n), 256, 0, c->stream>>>(c->g, c->d_ctJobs + (c->ctRingPos - n));
}

DeviceScope::DeviceScope(const x265cu_ctx* c) : prev(-1), want(c ? c->cfg.device : -1)
{
    if (want < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != want) cudaSetDevice(want);
}

template <typename T> T* slotPtr(x265cu_ctx* c, int slot, size_t off) { return (T*)(c->slots[slot] + off); }
/* the per-lowres-block AQ scale the block kernels read: invQscaleFactor, or invQscaleFactor8x8 with qg-size 8 */
int* slotInvQ(x265cu_ctx* c, int slot) { return slotPtr<int>(c, slot, c->g.aqBlock == 8 ? c->lay.invQ8 : c->lay.invQ); }

/* Green contexts through cudaGetDriverEntryPoint: no link-time dependency on libcuda (the library must load on a box
 * without a driver, where x265cu_create then fails with X265CU_ERR_NO_DEVICE).  Returns false when the partitioning is
 * unavailable or refused; the caller then creates ordinary priority streams. */
bool makeGreenStreams(x265cu_ctx* c, int smallSMs, int prGreatest, int prLeast)
{
    typedef CUresult (*GetRes)(CUdevice, CUdevResource*, CUdevResourceType);
    typedef CUresult (*Split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
    typedef CUresult (*GenDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
    typedef CUresult (*GreenCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    typedef CUresult (*GreenStream)(CUstream*, CUgreenCtx, unsigned int, int);
    typedef CUresult (*DevGet)(CUdevice*, int);
    void *fGetRes = NULL, *fSplit = NULL, *fGen = NULL, *fCreate = NULL, *fStream = NULL, *fDev = NULL;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuDeviceGetDevResource", &fGetRes, cudaEnableDefault, &q) != cudaSuccess || !fGetRes ||
        cudaGetDriverEntryPoint("cuDevSmResourceSplitByCount", &fSplit, cudaEnableDefault, &q) != cudaSuccess || !fSplit ||
        cudaGetDriverEntryPoint("cuDevResourceGenerateDesc", &fGen, cudaEnableDefault, &q) != cudaSuccess || !fGen ||
        cudaGetDriverEntryPoint("cuGreenCtxCreate", &fCreate, cudaEnableDefault, &q) != cudaSuccess || !fCreate ||
        cudaGetDriverEntryPoint("cuGreenCtxStreamCreate", &fStream, cudaEnableDefault, &q) != cudaSuccess || !fStream ||
        cudaGetDriverEntryPoint("cuDeviceGet", &fDev, cudaEnableDefault, &q) != cudaSuccess || !fDev)
    { cudaGetLastError(); return false; }
    CUdevice dev;
    if (((DevGet)fDev)(&dev, c->cfg.device) != CUDA_SUCCESS) return false;
    CUdevResource all, small, rest;
    if (((GetRes)fGetRes)(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    unsigned int groups = 1;
    if (((Split)fSplit)(&small, &groups, &all, &rest, 0, (unsigned)smallSMs) != CUDA_SUCCESS || groups != 1) return false;
    if (small.sm.smCount < 8 || rest.sm.smCount < 64) return false;
    CUdevResourceDesc dSmall, dRest;
    if (((GenDesc)fGen)(&dSmall, &small, 1) != CUDA_SUCCESS || ((GenDesc)fGen)(&dRest, &rest, 1) != CUDA_SUCCESS) return false;
    CUgreenCtx gS = NULL, gL = NULL;
    if (((GreenCreate)fCreate)(&gS, dSmall, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    if (((GreenCreate)fCreate)(&gL, dRest, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    CUstream sMain = NULL, sMirror = NULL, lanes[LA_NUM_LANES];
    if (((GreenStream)fStream)(&sMain, gS, CU_STREAM_NON_BLOCKING, prGreatest) != CUDA_SUCCESS ||
        ((GreenStream)fStream)(&sMirror, gS, CU_STREAM_NON_BLOCKING, prGreatest) != CUDA_SUCCESS) return false;
    for (int i = 0; i < LA_NUM_LANES; i++)
        if (((GreenStream)fStream)(&lanes[i], gL, CU_STREAM_NON_BLOCKING, prLeast) != CUDA_SUCCESS) return false;
    CUstream sGather = NULL;
    if (((GreenStream)fStream)(&sGather, gS, CU_STREAM_NON_BLOCKING, prGreatest) != CUDA_SUCCESS) return false;
    c->stream = (cudaStream_t)sMain; c->mirrorStream = (cudaStream_t)sMirror; c->gatherStream = (cudaStream_t)sGather;
    for (int i = 0; i < LA_NUM_LANES; i++) c->lanes[i] = (cudaStream_t)lanes[i];
    c->greenSmall = gS; c->greenLarge = gL;
    c->greenSmallSMs = (int)small.sm.smCount; c->greenLargeSMs = (int)rest.sm.smCount;
    return true;
}

int ensureDev(x265cu_ctx* c, char** p, size_t* cap, size_t need)
{
    if (*cap >= need) return X265CU_OK;
    if (*p) { cudaStreamSynchronize(c->stream); cudaFree(*p); *p = NULL; *cap = 0; }
    size_t n = alignUp(need * 2, 4096);
    CK(cudaMalloc((void**)p, n));
    *cap = n;
    return X265CU_OK;
}

bool slotOk(const x265cu_ctx* c, int s) { return s >= 0 && s < (int)c->slots.size(); }

/* ---------------------------------------------------------------- small results without the copy engine
 * The scalars the host waits for (frame statistics, cost sums, skip flags, recalculated scores) are a few bytes
 * each.  As cudaMemcpyAsync they queue on the copy engine behind whatever 16 MB picture chunk is in flight (measured:
 * ~0.3 ms per call while pictures upload), so kernels write them straight into mapped pinned host memory instead and
 * the host only waits for the event / stream. */
__global__ void publish_kernel(const unsigned* __restrict__ src, unsigned* __restrict__ dstMapped, int nWords)
{
    for (int i = threadIdx.x; i < nWords; i += blockDim.x) dstMapped[i] = src[i];
}

/* out[i] = wordsEach words read from srcs[i]; srcs and out live in mapped host memory */
__global__ void gather_small_kernel(const unsigned* const* __restrict__ srcs, int wordsEach, unsigned* __restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned* s = srcs[i];
    for (int w = 0; w < wordsEach; w++) out[(size_t)i * wordsEach + w] = __ldcg(s + w);
}

int ensureMapped(x265cu_ctx* c, size_t need)
{
    if (c->mappedCap >= need) return X265CU_OK;
    if (c->h_mapped) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->gatherStream); cudaFreeHost(c->h_mapped); c->h_mapped = NULL; c->mappedCap = 0; }
    const size_t n = alignUp(need * 2, 4096);
    CK(cudaHostAlloc((void**)&c->h_mapped, n, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&c->d_mapped, c->h_mapped, 0));
    c->mappedCap = n;
    return X265CU_OK;
}

/* wordsEach words from each of srcs[0..n) into dst (host), through one kernel on the main stream; synchronises */
int gatherSmall(x265cu_ctx* c, const std::vector<const void*>& srcs, int wordsEach, void* dst)
{
    HostTimer ht(HT_GATHER);
    const size_t n = srcs.size();
    const size_t tabBytes = alignUp(n * sizeof(void*), 256), outBytes = n * wordsEach * sizeof(unsigned);
    int st = ensureMapped(c, tabBytes + outBytes);
    if (st) return st;
    memcpy(c->h_mapped, &srcs[0], n * sizeof(void*));
    gather_small_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->gatherStream>>>((const unsigned* const*)c->d_mapped, wordsEach,
                                                                                   (unsigned*)(c->d_mapped + tabBytes), (int)n);
    c->counters.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->gatherStream));
    memcpy(dst, c->h_mapped + tabBytes, outBytes);
    c->counters.d2h_bytes += outBytes;
    return X265CU_OK;
}

/* Small host <-> device transfers on the paths the GPU waits for go through a KERNEL that reads / writes mapped page-locked host
 * memory, never through cudaMemcpyAsync: the copy engines are FIFO across streams, and in an end-to-end run the H2D engine
 * is saturated by picture uploads (25 MB every 0.46 ms at 2160p main10) -- a 36 KB job array queued behind them held the
 * first search launch of a stream back by 60 ms (every picture of the window fill went first), and the 0.5 KB result of a
 * cost recalculation waited 0.4 ms behind a 21 MB plane mirror on the D2H engine.  Also zeroes `zeroWords` words at `zero`. */
struct StageSeg { const uint4* src; uint4* dst; unsigned n16; unsigned pad; };
struct StageTab { StageSeg s[3]; unsigned* zero; unsigned zeroWords; };
__global__ void __launch_bounds__(256) stage_kernel(StageTab t)
{
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
#pragma unroll
    for (int k = 0; k < 3; k++)
        for (unsigned i = tid; i < t.s[k].n16; i += nth) t.s[k].dst[i] = t.s[k].src[i];
    for (unsigned i = tid; i < t.zeroWords; i += nth) t.zero[i] = 0;
}

/* device-visible address of page-locked host memory allocated by this library */
static void* hostAlias(void* h)
{
    void* d = NULL;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { cudaGetLastError(); return NULL; }
    return d;
}

/* up to three (dst, src, bytes) copies -- bytes a multiple of 16, both ends 16-byte aligned; the host side of each must be
 * page-locked memory of this library -- and one zero fill, as one launch on `stream` */
int stageLaunch(x265cu_ctx* c, cudaStream_t stream, void* dst0, const void* src0, size_t bytes0, void* dst1, const void* src1, size_t bytes1,
                unsigned* zero, size_t zeroWords)
{
    /* X265CU_STAGE_MEMCPY=1: the same through the copy engines (ncu cannot replay a kernel that reads mapped host memory --
     * "Failed to prepare kernel for profiling" --, so the ncu launch lists under profiles/ are taken with this set) */
    static const bool viaMemcpy = getenv("X265CU_STAGE_MEMCPY") != NULL;
    if (viaMemcpy)
    {
        if (bytes0) CK(cudaMemcpyAsync(dst0, src0, bytes0, cudaMemcpyDefault, stream));
        if (bytes1) CK(cudaMemcpyAsync(dst1, src1, bytes1, cudaMemcpyDefault, stream));
        if (zeroWords) CK(cudaMemsetAsync(zero, 0, zeroWords * 4, stream));
        return X265CU_OK;
    }
    StageTab t;
    memset(&t, 0, sizeof(t));
    t.s[0].dst = (uint4*)dst0; t.s[0].src = (const uint4*)src0; t.s[0].n16 = (unsigned)(bytes0 / 16);
    t.s[1].dst = (uint4*)dst1; t.s[1].src = (const uint4*)src1; t.s[1].n16 = (unsigned)(bytes1 / 16);
    t.zero = zero; t.zeroWords = (unsigned)zeroWords;
    const size_t work = std::max(std::max(bytes0, bytes1) / 16, zeroWords);
    if (!work) return X265CU_OK;
    const unsigned grid = (unsigned)std::max((size_t)1, std::min((size_t)64, (work + 255) / 256));
    stage_kernel<<<grid, 256, 0, stream>>>(t);
    c->counters.kernel_launches++;
    CK(cudaGetLastError());
    return X265CU_OK;
}

/* ---------------------------------------------------------------- batches */

Batch* batchOf(x265cu_ctx* c, long long id)
{
    if (id < 0) return NULL;
    Batch* b = &c->batches[id % LA_NUM_BATCHES];
    return b->id == id ? b : NULL;      /* NULL: the object was reused, i.e. batch `id` finished long ago */
}

int batchStage(x265cu_ctx* c, Batch* b, size_t bytes, char** h, char** d);

/* gather / scatter of the exchanged stores: one launch moves every segment (a cudaMemcpyAsync each cost more host
 * time than the whole batch).  Pointers are 256-byte aligned, lengths multiples of 4 bytes. */
struct SegCopy { const char* src; char* dst; unsigned long long bytes; };

__global__ void __launch_bounds__(256) copy_segments_kernel(const SegCopy* __restrict__ segs)
{
    const SegCopy s = segs[blockIdx.y];
    const unsigned long long n16 = s.bytes >> 4;
    const uint4* src = (const uint4*)s.src; uint4* dst = (uint4*)s.dst;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (unsigned long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
    if (blockIdx.x == 0)
    {
        const unsigned long long n4 = (s.bytes & 15) >> 2;
        if (threadIdx.x < n4)
            ((unsigned*)(s.dst + (n16 << 4)))[threadIdx.x] = ((const unsigned*)(s.src + (n16 << 4)))[threadIdx.x];
    }
}

/* exchange buffers come from a per-context pool: a batch borrows one per root and gives them back when it is done */
int xbufTake(x265cu_ctx* c, size_t need, char** p, size_t* cap)
{
    int best = -1;
    for (size_t i = 0; i < c->xpool.size(); i++)
        if (c->xpool[i].second >= need && (best < 0 || c->xpool[i].second < c->xpool[best].second)) best = (int)i;
    if (best >= 0)
    {
        *p = c->xpool[best].first; *cap = c->xpool[best].second;
        c->xpool.erase(c->xpool.begin() + best);
        return X265CU_OK;
    }
    *cap = alignUp(need * 3 / 2, 1 << 20);
    CK(cudaMalloc((void**)p, *cap));
    return X265CU_OK;
}
void xbufGiveBack(x265cu_ctx* c, Batch* b)
{
    for (int r = 0; r < LA_MAX_RANKS; r++)
        if (b->xbuf[r]) { c->xpool.push_back(std::make_pair(b->xbuf[r], b->xcap[r])); b->xbuf[r] = NULL; b->xcap[r] = 0; }
}

/* sharded stream: every store written in this batch travels from its owner to all the other ranks */
int exchangeBatch(x265cu_ctx* c, Batch* b)
{
    if (c->nranks <= 1 || b->segs.empty()) { b->segs.clear(); return X265CU_OK; }
    if (!c->exchange) { snprintf(c->err, sizeof(c->err), "sharded stream without an exchange callback"); return X265CU_ERR_BAD_ARG; }
    uint64_t total[LA_MAX_RANKS] = { 0 };
    for (size_t i = 0; i < b->segs.size(); i++) total[b->segs[i].root] += alignUp(b->segs[i].bytes, 256);
    for (int r = 0; r < c->nranks; r++)
        if (total[r] && !b->xbuf[r])
        {
            int st = xbufTake(c, total[r], &b->xbuf[r], &b->xcap[r]);
            if (st) return st;
        }
    /* copy tables: [0, nPack) owner -> exchange buffer, [nPack, n) exchange buffer -> slot on the other ranks */
    const size_t nseg = b->segs.size();
    char *hst, *dst;
    int st = batchStage(c, b, nseg * sizeof(SegCopy), &hst, &dst);
    if (st) return st;
    SegCopy* tab = (SegCopy*)hst;
    size_t nPack = 0;
    for (size_t i = 0; i < nseg; i++) nPack += b->segs[i].root == c->rank;
    size_t iPack = 0, iUnpack = nPack;
    size_t off[LA_MAX_RANKS] = { 0 };
    for (size_t i = 0; i < nseg; i++)
    {
        const Batch::Seg& s = b->segs[i];
        char* x = b->xbuf[s.root] + off[s.root];
        if (s.root == c->rank) { SegCopy sc = { s.ptr, x, s.bytes }; tab[iPack++] = sc; }
        else { SegCopy sc = { x, s.ptr, s.bytes }; tab[iUnpack++] = sc; }
        off[s.root] += alignUp(s.bytes, 256);
    }
    CK(cudaMemcpyAsync(dst, hst, nseg * sizeof(SegCopy), cudaMemcpyHostToDevice, b->stream));
    if (nPack)
    {
        copy_segments_kernel<<<dim3(16, (unsigned)nPack), 256, 0, b->stream>>>((const SegCopy*)dst);
        c->counters.kernel_launches++;
    }
    void* bufs[LA_MAX_RANKS];
    for (int r = 0; r < LA_MAX_RANKS; r++) bufs[r] = b->xbuf[r];
    if (c->exchange(c->exchangeUser, bufs, total, c->nranks, (void*)b->stream) != 0)
    { snprintf(c->err, sizeof(c->err), "exchange callback failed"); return X265CU_ERR_CUDA; }
    if (nseg > nPack)
    {
        copy_segments_kernel<<<dim3(16, (unsigned)(nseg - nPack)), 256, 0, b->stream>>>((const SegCopy*)dst + nPack);
        c->counters.kernel_launches++;
    }
    CK(cudaGetLastError());
    b->segs.clear();
    return X265CU_OK;
}

int endBatch(x265cu_ctx* c)
{
    if (!c->cur) return X265CU_OK;
    HostTimer ht(HT_BATCH_END);
    Batch* b = c->cur;
    c->cur = NULL;
    b->open = false;
    int st = exchangeBatch(c, b);
    if (st) return st;
    CK(cudaEventRecord(b->done, b->stream));
    return X265CU_OK;
}

int beginBatch(x265cu_ctx* c)
{
    HostTimer ht(HT_BATCH_BEGIN);
    int st = endBatch(c);
    if (st) return st;
    if (c->nranks > 1)
        for (int i = 0; i < LA_NUM_BATCHES; i++)
        {
            Batch& o = c->batches[i];
            bool holds = false;
            for (int r = 0; r < c->nranks; r++) holds |= o.xbuf[r] != NULL;
            if (holds && o.id >= 0 && !o.open && cudaEventQuery(o.done) == cudaSuccess)
                xbufGiveBack(c, &o);
        }
    Batch* b = &c->batches[c->nextBatch % LA_NUM_BATCHES];
    if (b->id >= 0) CK(cudaEventSynchronize(b->done));       /* blocks only with LA_NUM_BATCHES batches in flight */
    xbufGiveBack(c, b);
    for (size_t i = 0; i < b->retiredDev.size(); i++) cudaFree(b->retiredDev[i]);
    for (size_t i = 0; i < b->retiredHost.size(); i++) cudaFreeHost(b->retiredHost[i]);
    b->retiredDev.clear(); b->retiredHost.clear(); b->waited.clear(); b->waitedDone.clear();
    b->id = c->nextBatch++;
    b->stream = c->lanes[b->id % LA_NUM_LANES];
    b->stageUsed = 0; b->syncUsed = 0; b->open = true;
    /* the batch reads planes / intra costs / AQ factors of every frame uploaded so far */
    for (int i = 0; i < LA_NUM_PRE; i++)
    {
        CK(cudaEventRecord(b->begun[i], c->preStreams[i]));
        CK(cudaStreamWaitEvent(b->stream, b->begun[i], 0));
    }
    c->cur = b;
    return X265CU_OK;
}

/* room for `bytes` of job array in the batch's pinned + device staging */
int batchStage(x265cu_ctx* c, Batch* b, size_t bytes, char** h, char** d)
{
    bytes = alignUp(bytes, 256);
    if (b->stageUsed + bytes > b->stageCap)
    {
        if (b->h_stage) { b->retiredHost.push_back(b->h_stage); b->retiredDev.push_back(b->d_stage); }
        b->h_stage = NULL; b->d_stage = NULL; b->stageUsed = 0;
        b->stageCap = alignUp(std::max(bytes * 2, b->stageCap * 2), 4096);
        CK(cudaHostAlloc((void**)&b->h_stage, b->stageCap, cudaHostAllocMapped));
        CK(cudaMalloc((void**)&b->d_stage, b->stageCap));
    }
    *h = b->h_stage + b->stageUsed; *d = b->d_stage + b->stageUsed;
    b->stageUsed += bytes;
    return X265CU_OK;
}

int batchSync(x265cu_ctx* c, Batch* b, size_t ints, int** d)
{
    const size_t bytes = alignUp(ints * sizeof(int), 256);
    if (b->syncUsed + bytes > b->syncCap)
    {
        if (b->d_sync) b->retiredDev.push_back(b->d_sync);
        b->d_sync = NULL; b->syncUsed = 0;
        b->syncCap = alignUp(std::max(bytes * 2, b->syncCap * 2), 4096);
        CK(cudaMalloc((void**)&b->d_sync, b->syncCap));
    }
    *d = (int*)((char*)b->d_sync + b->syncUsed);
    b->syncUsed += bytes;
    return X265CU_OK;
}

/* the batch's stream waits for the searches of batch `id` (another lane) */
int batchWaitSearches(x265cu_ctx* c, Batch* b, long long id)
{
    if (id < 0 || id == b->id) return X265CU_OK;
    if (std::find(b->waited.begin(), b->waited.end(), id) != b->waited.end()) return X265CU_OK;
    b->waited.push_back(id);
    Batch* w = batchOf(c, id);
    if (w) CK(cudaStreamWaitEvent(b->stream, w->searchDone, 0));
    return X265CU_OK;
}

/* The batch is about to overwrite a store that batch `id` wrote (a search redone with weights after it had been enqueued
 * assuming none, Lookahead::verifyWeights): everything of that batch -- its writers and the cost jobs reading the store --
 * goes first.  Rare (fades); the stores of a recycled slot are reset by the upload, which orders the tenants itself. */
int batchWaitDone(x265cu_ctx* c, Batch* b, long long id)
{
    if (id < 0 || id == b->id) return X265CU_OK;
    if (std::find(b->waitedDone.begin(), b->waitedDone.end(), id) != b->waitedDone.end()) return X265CU_OK;
    b->waitedDone.push_back(id);
    Batch* w = batchOf(c, id);
    if (w && !w->open) CK(cudaStreamWaitEvent(b->stream, w->done, 0));
    return X265CU_OK;
}

/* the main stream waits for batch `id` (its searches only, or all of it) */
int mainWaitBatch(x265cu_ctx* c, long long id, bool searchesOnly)
{
    Batch* w = batchOf(c, id);
    if (!w) return X265CU_OK;
    if (w->open) { int st = endBatch(c); if (st) return st; }
    CK(cudaStreamWaitEvent(c->stream, searchesOnly ? w->searchDone : w->done, 0));
    return X265CU_OK;
}
/* `st` waits for batch `id` (its searches only, or all of it) */
int streamWaitBatch(x265cu_ctx* c, cudaStream_t st, long long id, bool searchesOnly)
{
    Batch* w = batchOf(c, id);
    if (!w) return X265CU_OK;
    if (w->open) { int rc = endBatch(c); if (rc) return rc; }
    CK(cudaStreamWaitEvent(st, searchesOnly ? w->searchDone : w->done, 0));
    return X265CU_OK;
}

int mainWaitMv(x265cu_ctx* c, int slot, int store)
{
    /* sharded stream: the MVs of a frame another rank owns arrive with the exchange at the end of the batch */
    return store < 0 ? X265CU_OK : mainWaitBatch(c, c->mvWriter[(size_t)slot * c->geom.n_mv_stores + store], c->nranks <= 1);
}
int mainWaitCost(x265cu_ctx* c, int slot, int store)
{
    return store < 2 ? X265CU_OK : mainWaitBatch(c, c->costWriter[(size_t)slot * c->geom.n_cost_stores + store], false);
}

/* main-stream work is about to read what the pre-lookahead of `slot` produced (planes, intra costs, AQ arrays) */
int mainWaitPre(x265cu_ctx* c, int slot)
{
    c->slotMainTouched[slot] = 1;
    CK(cudaStreamWaitEvent(c->stream, c->slotConsumed[slot], 0));
    return X265CU_OK;
}

void touchSlot(x265cu_ctx* c, int slot, long long id)
{
    std::vector<long long>& u = c->slotUsers[slot];
    if (std::find(u.begin(), u.end(), id) == u.end()) u.push_back(id);
}

char* mvStorePtr(x265cu_ctx* c, int slot, int store) { return c->slots[slot] + c->lay.mvStores + (size_t)store * c->lay.mvStoreStride; }
char* costStorePtr(x265cu_ctx* c, int slot, int store) { return c->slots[slot] + c->lay.costStores + (size_t)store * c->lay.costStoreStride; }

/* ---------------------------------------------------------------- typed implementations */

template <typename P>
int uploadT(x265cu_ctx* c, int slot, const void* y, const void* u, const void* v, int sy, int sc)
{
    HostTimer ht(HT_UPLOAD);
    const Geom& g = c->g;
    const SlotLayout& L = c->lay;
    P* dY = slotPtr<P>(c, slot, L.srcY);
    P* dU = slotPtr<P>(c, slot, L.srcU);
    P* dV = slotPtr<P>(c, slot, L.srcV);
    /* the copy may not overwrite the staging planes while the previous tenant's K1/K2 still read them */
    CK(cudaStreamWaitEvent(c->copyStream, c->slotConsumed[slot], 0));
    /* cudaMemcpyDefault: the picture may live in host memory (pageable or pinned) or already in HBM */
    /* a contiguous plane goes as one linear copy (one DMA descriptor instead of one per row) */
    if (sy == g.picW && g.srcPitch == g.picW) CK(cudaMemcpyAsync(dY, y, (size_t)g.picW * g.picH * sizeof(P), cudaMemcpyDefault, c->copyStream));
    else CK(cudaMemcpy2DAsync(dY, g.srcPitch * sizeof(P), y, (size_t)sy * sizeof(P), g.picW * sizeof(P), g.picH, cudaMemcpyDefault, c->copyStream));
    c->counters.h2d_bytes += (uint64_t)g.picW * g.picH * sizeof(P);
    const bool chroma = u && v;
    if (chroma && c->cfg.need_aq)
    {
        if (sc == g.cW && g.srcPitchC == g.cW)
        {
            CK(cudaMemcpyAsync(dU, u, (size_t)g.cW * g.cH * sizeof(P), cudaMemcpyDefault, c->copyStream));
            CK(cudaMemcpyAsync(dV, v, (size_t)g.cW * g.cH * sizeof(P), cudaMemcpyDefault, c->copyStream));
        }
        else
        {
            CK(cudaMemcpy2DAsync(dU, g.srcPitchC * sizeof(P), u, (size_t)sc * sizeof(P), g.cW * sizeof(P), g.cH, cudaMemcpyDefault, c->copyStream));
            CK(cudaMemcpy2DAsync(dV, g.srcPitchC * sizeof(P), v, (size_t)sc * sizeof(P), g.cW * sizeof(P), g.cH, cudaMemcpyDefault, c->copyStream));
        }
        c->counters.h2d_bytes += 2ull * g.cW * g.cH * sizeof(P);
    }
    CK(cudaEventRecord(c->slotCopied[slot], c->copyStream));
    const cudaStream_t ps = c->preStreams[c->preSeq++ % LA_NUM_PRE];
    CK(cudaStreamWaitEvent(ps, c->slotCopied[slot], 0));
    /* K1-K3 overwrite the slot: every batch that still reads or writes its previous tenant goes first (the copy
     * above only touches the staging planes, which batches never read, so it is not held back) */
    {
        std::vector<long long>& users = c->slotUsers[slot];
        for (size_t i = 0; i < users.size(); i++)
        {
            Batch* w = batchOf(c, users[i]);
            if (!w) continue;
            if (w->open) { int st = endBatch(c); if (st) return st; }
            CK(cudaStreamWaitEvent(ps, w->done, 0));
        }
        users.clear();
        std::fill(c->mvWriter.begin() + (size_t)slot * c->geom.n_mv_stores, c->mvWriter.begin() + (size_t)(slot + 1) * c->geom.n_mv_stores, -1LL);
        std::fill(c->costWriter.begin() + (size_t)slot * c->geom.n_cost_stores, c->costWriter.begin() + (size_t)(slot + 1) * c->geom.n_cost_stores, -1LL);
    }
    if (c->slotMirrorTouched[slot])
    {
        CK(cudaStreamWaitEvent(ps, c->slotMirrored[slot], 0));
        c->slotMirrorTouched[slot] = 0;
    }
    c->slotRecalcStore[slot] = -1;
    c->slotMainWroteSet[slot] = 0;
    if (c->slotMainTouched[slot])
    {
        /* ... and whatever the main stream (cuTree, recalc, mirrors) still reads of it */
        CK(cudaEventRecord(c->mainMark, c->stream));
        CK(cudaStreamWaitEvent(ps, c->mainMark, 0));
        c->slotMainTouched[slot] = 0;
    }
    /* stats + rowSatds00 start at zero */
    CK(cudaMemsetAsync(c->slots[slot] + L.rowSatds00, 0, L.stats + sizeof(FrameStatsDev) - L.rowSatds00, ps));
    P* planes = slotPtr<P>(c, slot, L.planes);
    FrameStatsDev* stats = slotPtr<FrameStatsDev>(c, slot, L.stats);
    int* invQ = slotInvQ(c, slot);
    const bool qg8 = g.aqBlock == 8;
    unsigned* energy = slotPtr<unsigned>(c, slot, L.energy);
    {
        /* K1 + K2a fused: lowres planes and (qg-size > 8) the 16x16 AC energies / weightp sums in one pass over the luma */
        Prof pr(c, X265CU_K_LOWRES, 2, ps);
        /* persistent CTAs: LA_LR_CTAS_PER_SM per SM, each walking groups of LA_LR_TILES tiles with a ring of staged copies */
        const int groups = (g.bw + LA_LR_TILES - 1) / LA_LR_TILES * g.bh;
        const int grid = std::min(groups, c->numSMs * LA_LR_CTAS_PER_SM);
        lowres_fused_kernel<P><<<grid, 32 * LA_LR_TILES, 0, ps>>>(g, dY, chroma ? dU : NULL, chroma ? dV : NULL, planes, energy, stats,
                                                                  c->cfg.need_aq && !qg8);
        const long long tileRows = 4LL * g.tpr * g.planeLines;
        extend_border_kernel<P><<<(unsigned)((tileRows + 255) / 256), 256, 0, ps>>>(g, planes);
    }
    if (c->cfg.hme)
    {
        /* the 1/16-resolution planes of HME level 0 from the finished lowresPlane[0] (lowres.cpp:378-388) */
        const Geom& g4 = c->g4;
        Prof pr(c, X265CU_K_LOWRES, 2, ps);
        P* planes4 = slotPtr<P>(c, slot, L.planes4);
        const long long n4 = (long long)((g4.w + 7) & ~7) * g4.h;
        lowerres_kernel<P><<<(unsigned)((n4 + 255) / 256), 256, 0, ps>>>(g, g4, planes, planes4);
        const long long tileRows4 = 4LL * g4.tpr * g4.planeLines;
        extend_border_kernel<P><<<(unsigned)((tileRows4 + 255) / 256), 256, 0, ps>>>(g4, planes4);
    }
    if (c->cfg.need_aq)
    {
        const bool twoPass = c->cfg.aq_mode >= 2 && c->cfg.aq_strength != 0;
        const bool edgeAq = c->cfg.aq_mode > 3 && c->cfg.aq_strength != 0;
        unsigned* edge = edgeAq ? slotPtr<unsigned>(c, slot, L.edge) : NULL;
        Prof pr(c, X265CU_K_AQ, 1 + 2 * twoPass + 2 * qg8 + edgeAq, ps);
        double* qpCuTree = slotPtr<double>(c, slot, L.qpCuTree);
        double* sums = slotPtr<double>(c, slot, L.aqSums);
        if (qg8)
            aq_energy8_kernel<P><<<(g.aqW * g.aqH + 31) / 32, 256, 0, ps>>>(g, dY, chroma ? dU : NULL, chroma ? dV : NULL, energy, stats);
        if (edgeAq)
            aq_edge_kernel<P><<<((g.picW + 15) >> 4) * ((g.picH + 15) >> 4), 256, 0, ps>>>(g, dY, edge, stats);
        if (twoPass)
        {
            aq_pow_kernel<<<LA_AQ_CTAS, LA_AQ_THREADS, 0, ps>>>(g, energy, edge, qpCuTree);
            aq_mean_kernel<<<1, 32, 0, ps>>>(g, qpCuTree, sums);
        }
        aq_finish_kernel<<<LA_AQ_CTAS, LA_AQ_THREADS, 0, ps>>>(g, energy, c->cfg.aq_mode, c->cfg.aq_strength, c->cfg.need_wp_stats,
                                                                      c->cfg.fade_stats, sums, slotPtr<double>(c, slot, L.qpAq), qpCuTree,
                                                                      slotPtr<int>(c, slot, L.invQ), stats, edge);
        if (qg8)
            aq_invq8x8_kernel<<<(g.ncu + 255) / 256, 256, 0, ps>>>(g, slotPtr<int>(c, slot, L.invQ), invQ);
    }
    {
        Prof pr(c, X265CU_K_INTRA, 1, ps);
        intra_kernel<P><<<(g.ncu + 15) / 16, 128, 0, ps>>>(g, planes, c->cfg.need_aq ? invQ : NULL,
                                                                  slotPtr<int>(c, slot, L.intraCost), slotPtr<unsigned char>(c, slot, L.intraMode),
                                                                  slotPtr<unsigned short>(c, slot, L.lowresCosts00),
                                                                  slotPtr<int>(c, slot, L.rowSatds00), stats);
    }
    if (c->cfg.hist_stats)
    {
        /* --hist-scenecut picture statistics (8-bit only, checked at create) */
        HistAcc* acc = c->d_histAcc + slot;
        int hst = stageLaunch(c, ps, NULL, NULL, 0, NULL, NULL, 0, (unsigned*)acc, sizeof(HistAcc) / 4);
        if (hst) return hst;
        Prof pr(c, X265CU_K_AQ, 4, ps);
        const uint8_t* y8 = (const uint8_t*)dY; const uint8_t* u8 = (const uint8_t*)dU; const uint8_t* v8 = (const uint8_t*)dV;
        hist_luma_kernel<<<dim3(16, LA_HIST_CHUNKS), 256, 0, ps>>>(g, (const uint8_t*)planes, acc);
        hist_chroma_kernel<<<dim3(16, 2), 256, 0, ps>>>(g, u8, v8, acc);
        hist_var_kernel<<<(g.picH + 7) / 8 + 2 * (((g.picH >> 1) + 3) / 4), 256, 0, ps>>>(g, y8, u8, v8, acc);
        hist_finish_kernel<<<1, 256, 0, ps>>>(g, acc, c->d_hist + slot);
    }
    publish_kernel<<<1, 32, 0, ps>>>((const unsigned*)stats, (unsigned*)(c->d_slotStats + slot), (int)(sizeof(FrameStatsDev) / 4));
    c->counters.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->slotConsumed[slot], ps));
    return X265CU_OK;
}

template <typename P>
int weightPlanes(x265cu_ctx* c, cudaStream_t stream, const P* src, P* dst, int nPlanes, int scale, int denom, int offsetIn)
{
    const int correction = 14 - c->g.depth;
    const int offset = offsetIn << (c->g.depth - 8);
    const int round = (denom ? 1 << (denom - 1) : 0) << correction;
    const int shift = denom + correction;
    const long long n = c->g.planeSize * nPlanes;
    Prof pr(c, X265CU_K_WEIGHT, 1, stream);
    weight_planes_kernel<P><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, n, scale, round, shift, offset, correction,
                                                                                (1 << c->g.depth) - 1);
    CK(cudaGetLastError());
    return X265CU_OK;
}

int ensureScratch(x265cu_ctx* c, std::vector<char*>& scratch, cudaStream_t stream, size_t count)
{
    while (scratch.size() < count)
    {
        char* p = NULL;
        CK(cudaMalloc((void**)&p, (size_t)(4 * c->g.planeSize) * c->bpp + 256));
        CK(cudaMemsetAsync(p, 0, (size_t)(4 * c->g.planeSize) * c->bpp + 256, stream));
        scratch.push_back(p);
    }
    return X265CU_OK;
}

template <typename P> int searchBatchHmeT(x265cu_ctx* c, const x265cu_search_job* jobs, int n);

template <typename P>
int searchBatchT(x265cu_ctx* c, const x265cu_search_job* jobs, int n)
{
    if (c->cfg.hme) return searchBatchHmeT<P>(c, jobs, n);
    HostTimer ht(HT_SEARCH_ENQ);
    const Geom& g = c->g;
    const SlotLayout& L = c->lay;
    const bool implicit = !c->cur;
    if (implicit) { int st = beginBatch(c); if (st) return st; }
    Batch* b = c->cur;
    char *hst, *dst;
    int st = batchStage(c, b, n * sizeof(SearchJobDev<P>), &hst, &dst);
    if (st) return st;
    SearchJobDev<P>* dev = (SearchJobDev<P>*)hst;
    /* weighted references: one scratch set per distinct (ref, weight) in this call */
    std::map<std::vector<int>, int> wmap;
    const int nAll = n;
    int nm = 0;         /* jobs this rank computes (all of them unless the stream is sharded) */
    for (int i = 0; i < nAll; i++)
    {
        const x265cu_search_job& j = jobs[i];
        if (!slotOk(c, j.fenc_slot) || !slotOk(c, j.ref_slot) || j.store < 0 || j.store >= c->geom.n_mv_stores ||
            j.cond_store >= c->geom.n_mv_stores)
        { snprintf(c->err, sizeof(c->err), "search job %d: bad slot/store", i); return X265CU_ERR_BAD_ARG; }
        char* ms = mvStorePtr(c, j.fenc_slot, j.store);
        touchSlot(c, j.fenc_slot, b->id); touchSlot(c, j.ref_slot, b->id);
        st = batchWaitDone(c, b, c->mvWriter[(size_t)j.fenc_slot * c->geom.n_mv_stores + j.store]);
        if (st) return st;
        c->mvWriter[(size_t)j.fenc_slot * c->geom.n_mv_stores + j.store] = b->id;
        if (c->nranks > 1)
        {
            Batch::Seg sg = { ms, (size_t)g.ncu * 8 + 4, c->slotOwner[j.fenc_slot] };
            b->segs.push_back(sg);
            if (sg.root != c->rank) continue;
        }
        const P* refBuf = slotPtr<P>(c, j.ref_slot, L.planes);
        if (j.weighted)
        {
            std::vector<int> key(4);
            key[0] = j.ref_slot; key[1] = j.w_scale; key[2] = j.w_denom; key[3] = j.w_offset;
            std::map<std::vector<int>, int>::iterator it = wmap.find(key);
            int idx;
            if (it == wmap.end())
            {
                idx = (int)wmap.size();
                wmap[key] = idx;
                st = ensureScratch(c, b->weightScratch, b->stream, idx + 1);
                if (st) return st;
                st = weightPlanes<P>(c, b->stream, refBuf, (P*)b->weightScratch[idx], 4, j.w_scale, j.w_denom, j.w_offset);
                if (st) return st;
            }
            else
                idx = it->second;
            refBuf = (const P*)b->weightScratch[idx];
        }
        SearchJobDev<P>& D = dev[nm++];
        D.fenc0 = slotPtr<P>(c, j.fenc_slot, L.planes);
        D.ref0 = refBuf;
        D.mvOut = (int*)ms;
        D.costOut = (int*)ms + g.ncu;
        D.flagOut = (int*)ms + 2 * g.ncu;
        D.cond = NULL;
        if (j.cond_store >= 0)
        {
            D.cond = (const int*)mvStorePtr(c, j.fenc_slot, j.cond_store) + 2 * g.ncu;
            st = batchWaitSearches(c, b, c->mvWriter[(size_t)j.fenc_slot * c->geom.n_mv_stores + j.cond_store]);
            if (st) return st;
        }
        D.bidir = j.bidir_ctx;
        D.sliced = j.sliced;
    }
    n = nm;
    if (!n)
    {
        CK(cudaEventRecord(b->searchDone, b->stream));
        return implicit ? endBatch(c) : X265CU_OK;
    }
    const int stripRows = 32 / c->searchLanes;
    const int nstrips = (g.bh + stripRows - 1) / stripRows;      /* strips of 4 or 8 block rows, one warp each */
    c->searchEnq += n;
    int* dsync;
    st = batchSync(c, b, (size_t)(1 + n * nstrips), &dsync);
    if (st) return st;
    {
        void* hAlias = hostAlias(hst);
        if (!hAlias) { snprintf(c->err, sizeof(c->err), "job staging memory is not device-visible"); return X265CU_ERR_CUDA; }
        st = stageLaunch(c, b->stream, dst, hAlias, alignUp(n * sizeof(SearchJobDev<P>), 16), NULL, NULL, 0, (unsigned*)dsync, (size_t)(1 + n * nstrips));
        if (st) return st;
    }
    if (n >= 64 && c->bigSearch[1] != b->id)
    {
        /* Large search launches of different batches run on different lanes with equal priority; left alone, the launches of
         * several queued batches share the GPU evenly and all finish late together -- the host, which needs the OLDEST batch
         * first, then waits for the sum of them.  At most two overlap (the second ramps up while the first drains): a large
         * launch starts after the one two before it has finished.  Small (on-demand) launches are never held back. */
        Batch* w = batchOf(c, c->bigSearch[0]);
        if (w && !w->open) CK(cudaStreamWaitEvent(b->stream, w->searchDone, 0));
        c->bigSearch[0] = c->bigSearch[1]; c->bigSearch[1] = b->id;
    }
    {
        Prof pr(c, X265CU_K_SEARCH, 1, b->stream);
        /* workers per job: a job's strips start 8 steps apart and run ~bw steps, so beyond ~bw/8 of them some only
         * sit resident waiting for their turn; that matters when the launch is too small to oversubscribe the GPU */
        const int workers = c->searchOneShot ? nstrips
                                             : std::max(1, std::min(nstrips, c->searchWorkers > 0 ? c->searchWorkers : (nstrips + 1) / 2));
        /* searchSmem: bytes of (unused) dynamic shared memory per one-warp CTA -- caps how many search warps an SM holds,
         * so that the short high-priority kernels (pre-lookahead, cuTree, recalc) always find room beside them */
        if (c->searchLanes == 8)
            search_kernel<P, 8><<<n * workers, 32, c->searchSmem, b->stream>>>(g, (const SearchJobDev<P>*)dst, nstrips, n,
                                                           c->d_mvcost + c->cfg.mvcost_half, dsync, dsync + 1, c->d_executed, c->searchOneShot);
        else
            search_kernel<P, 4><<<n * workers, 32, c->searchSmem, b->stream>>>(g, (const SearchJobDev<P>*)dst, nstrips, n,
                                                           c->d_mvcost + c->cfg.mvcost_half, dsync, dsync + 1, c->d_executed, c->searchOneShot);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(b->searchDone, b->stream));
    if (implicit) return endBatch(c);
    return X265CU_OK;
}

/* --hme: every search job is two launches of search_hme_kernel on the batch's stream -- level 0 on the 1/16-resolution planes
 * into the slot's level-0 store of the same index, then level 1 (the lowres search proper) reading it.  Bookkeeping (writer
 * table, slot users, shard segments, weighted copies, conditions) is that of searchBatchT. */
template <typename P>
int searchBatchHmeT(x265cu_ctx* c, const x265cu_search_job* jobs, int n)
{
    HostTimer ht(HT_SEARCH_ENQ);
    const Geom& g = c->g;
    const Geom& g4 = c->g4;
    const SlotLayout& L = c->lay;
    const bool implicit = !c->cur;
    if (implicit) { int st = beginBatch(c); if (st) return st; }
    Batch* b = c->cur;
    char *hst, *dst;
    int st = batchStage(c, b, 2 * alignUp((size_t)n * sizeof(SearchJobHme<P>), 16), &hst, &dst);
    if (st) return st;
    std::vector<SearchJobHme<P> > lv0, lv1;
    std::map<std::vector<int>, int> wmap;
    for (int i = 0; i < n; i++)
    {
        const x265cu_search_job& j = jobs[i];
        if (!slotOk(c, j.fenc_slot) || !slotOk(c, j.ref_slot) || j.store < 0 || j.store >= c->geom.n_mv_stores ||
            j.cond_store >= c->geom.n_mv_stores || j.sliced)
        { snprintf(c->err, sizeof(c->err), "search job %d: bad slot/store", i); return X265CU_ERR_BAD_ARG; }
        char* ms = mvStorePtr(c, j.fenc_slot, j.store);
        touchSlot(c, j.fenc_slot, b->id); touchSlot(c, j.ref_slot, b->id);
        st = batchWaitDone(c, b, c->mvWriter[(size_t)j.fenc_slot * c->geom.n_mv_stores + j.store]);
        if (st) return st;
        c->mvWriter[(size_t)j.fenc_slot * c->geom.n_mv_stores + j.store] = b->id;
        if (c->nranks > 1)
        {
            Batch::Seg sg = { ms, (size_t)g.ncu * 8 + 4, c->slotOwner[j.fenc_slot] };
            b->segs.push_back(sg);
            if (sg.root != c->rank) continue;
        }
        const P* refBuf = slotPtr<P>(c, j.ref_slot, L.planes);
        if (j.weighted)
        {
            std::vector<int> key(4);
            key[0] = j.ref_slot; key[1] = j.w_scale; key[2] = j.w_denom; key[3] = j.w_offset;
            std::map<std::vector<int>, int>::iterator it = wmap.find(key);
            int idx;
            if (it == wmap.end())
            {
                idx = (int)wmap.size();
                wmap[key] = idx;
                st = ensureScratch(c, b->weightScratch, b->stream, idx + 1);
                if (st) return st;
                st = weightPlanes<P>(c, b->stream, refBuf, (P*)b->weightScratch[idx], 4, j.w_scale, j.w_denom, j.w_offset);
                if (st) return st;
            }
            else
                idx = it->second;
            refBuf = (const P*)b->weightScratch[idx];
        }
        int* ms4 = (int*)(c->slots[j.fenc_slot] + L.mvStores4 + (size_t)j.store * L.mvStore4Stride);
        SearchJobHme<P> D0, D1;
        D0.fenc0 = slotPtr<P>(c, j.fenc_slot, L.planes4);
        D0.ref0 = slotPtr<P>(c, j.ref_slot, L.planes4);         /* never the weighted reference (slicetype.cpp:4083) */
        D0.mvOut = ms4; D0.costOut = ms4 + g4.ncu;
        D0.flagOut = (int*)ms + 2 * g.ncu;
        D0.cond = NULL;
        if (j.cond_store >= 0)
        {
            D0.cond = (const int*)mvStorePtr(c, j.fenc_slot, j.cond_store) + 2 * g.ncu;
            st = batchWaitSearches(c, b, c->mvWriter[(size_t)j.fenc_slot * c->geom.n_mv_stores + j.cond_store]);
            if (st) return st;
        }
        D0.hmeMv = NULL; D0.hmeCost = NULL;
        D0.bidir = j.bidir_ctx; D0.pad = 0;
        D1 = D0;
        D1.fenc0 = slotPtr<P>(c, j.fenc_slot, L.planes);
        D1.ref0 = refBuf;
        D1.mvOut = (int*)ms; D1.costOut = (int*)ms + g.ncu;
        D1.hmeMv = ms4; D1.hmeCost = ms4 + g4.ncu;
        lv0.push_back(D0); lv1.push_back(D1);
    }
    n = (int)lv0.size();
    if (!n)
    {
        CK(cudaEventRecord(b->searchDone, b->stream));
        return implicit ? endBatch(c) : X265CU_OK;
    }
    const size_t half = alignUp((size_t)n * sizeof(SearchJobHme<P>), 16);
    memcpy(hst, &lv0[0], (size_t)n * sizeof(SearchJobHme<P>));
    memcpy(hst + half, &lv1[0], (size_t)n * sizeof(SearchJobHme<P>));
    const int nstrips0 = (g4.bh + 3) / 4, nstrips1 = (g.bh + 3) / 4;
    c->searchEnq += n;
    int* dsync;
    const size_t sync0 = 1 + (size_t)n * nstrips0, sync1 = 1 + (size_t)n * nstrips1;
    st = batchSync(c, b, sync0 + sync1, &dsync);
    if (st) return st;
    {
        void* hAlias = hostAlias(hst);
        if (!hAlias) { snprintf(c->err, sizeof(c->err), "job staging memory is not device-visible"); return X265CU_ERR_CUDA; }
        st = stageLaunch(c, b->stream, dst, hAlias, half + alignUp((size_t)n * sizeof(SearchJobHme<P>), 16), NULL, NULL, 0, (unsigned*)dsync, sync0 + sync1);
        if (st) return st;
    }
    {
        Prof pr(c, X265CU_K_SEARCH, 2, b->stream);
        const unsigned short* mvc = c->d_mvcost + c->cfg.mvcost_half;
        search_hme_kernel<P><<<n * std::max(1, (nstrips0 + 1) / 2), 32, 0, b->stream>>>(g4, (const SearchJobHme<P>*)dst, nstrips0, n, mvc,
                                                                                         dsync, dsync + 1, c->d_executed, 0,
                                                                                         c->cfg.hme_search[0], c->cfg.hme_range[0]);
        search_hme_kernel<P><<<n * std::max(1, (nstrips1 + 1) / 2), 32, 0, b->stream>>>(g, (const SearchJobHme<P>*)(dst + half), nstrips1, n, mvc,
                                                                                         dsync + sync0, dsync + sync0 + 1, c->d_executed, 1,
                                                                                         c->cfg.hme_search[1], c->cfg.hme_range[1]);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(b->searchDone, b->stream));
    if (implicit) return endBatch(c);
    return X265CU_OK;
}

template <typename P>
int costBatchT(x265cu_ctx* c, const x265cu_cost_job* jobs, int n)
{
    HostTimer ht(HT_COST_ENQ);
    const Geom& g = c->g;
    const SlotLayout& L = c->lay;
    const bool implicit = !c->cur;
    if (implicit) { int st = beginBatch(c); if (st) return st; }
    Batch* b = c->cur;
    char *hst, *dst;
    int st = batchStage(c, b, n * sizeof(CostJobDev<P>), &hst, &dst);
    if (st) return st;
    /* P estimates (one thread per block) first, then B estimates (8 lanes per block): two launches with the
     * grid each needs */
    CostJobDev<P>* dev = (CostJobDev<P>*)hst;
    const int nAll = n;
    int nP = 0, nMine = 0;
    for (int i = 0; i < nAll; i++)
    {
        if (!slotOk(c, jobs[i].b_slot)) { snprintf(c->err, sizeof(c->err), "cost job %d: bad slot", i); return X265CU_ERR_BAD_ARG; }
        const bool mine = c->nranks <= 1 || c->slotOwner[jobs[i].b_slot] == c->rank;
        nMine += mine;
        nP += mine && jobs[i].l1_store < 0;
    }
    int iP = 0, iB = nP;
    const int nmv = c->geom.n_mv_stores;
    for (int i = 0; i < nAll; i++)
    {
        const x265cu_cost_job& j = jobs[i];
        if (!slotOk(c, j.b_slot) || !slotOk(c, j.p0_slot) || !slotOk(c, j.p1_slot) || j.out < 2 || j.out >= c->geom.n_cost_stores ||
            j.l0_store < 0 || j.l0_store >= nmv || j.l1_store >= nmv || j.cond_store >= nmv)
        { snprintf(c->err, sizeof(c->err), "cost job %d: bad slot/store", i); return X265CU_ERR_BAD_ARG; }
        char* cs = costStorePtr(c, j.b_slot, j.out);
        touchSlot(c, j.b_slot, b->id); touchSlot(c, j.p0_slot, b->id); touchSlot(c, j.p1_slot, b->id);
        st = batchWaitDone(c, b, c->costWriter[(size_t)j.b_slot * c->geom.n_cost_stores + j.out]);
        if (st) return st;
        c->costWriter[(size_t)j.b_slot * c->geom.n_cost_stores + j.out] = b->id;
        if (c->nranks > 1)
        {
            Batch::Seg sg = { cs, L.costResOff + sizeof(CostResultDev), c->slotOwner[j.b_slot] };
            b->segs.push_back(sg);
            if (sg.root != c->rank) continue;
        }
        CostJobDev<P>& d = dev[j.l1_store < 0 ? iP++ : iB++];
        d.fenc0 = slotPtr<P>(c, j.b_slot, L.planes);
        d.ref0 = slotPtr<P>(c, j.p0_slot, L.planes);
        d.ref1 = j.l1_store >= 0 ? slotPtr<P>(c, j.p1_slot, L.planes) : NULL;
        char* m0 = mvStorePtr(c, j.b_slot, j.l0_store);
        d.mv0 = (const int*)m0; d.cost0 = (const int*)m0 + g.ncu;
        st = batchWaitSearches(c, b, c->mvWriter[(size_t)j.b_slot * nmv + j.l0_store]);
        if (st) return st;
        if (j.l1_store >= 0)
        {
            char* m1 = mvStorePtr(c, j.b_slot, j.l1_store);
            d.mv1 = (const int*)m1; d.cost1 = (const int*)m1 + g.ncu;
            st = batchWaitSearches(c, b, c->mvWriter[(size_t)j.b_slot * nmv + j.l1_store]);
            if (st) return st;
        }
        else { d.mv1 = NULL; d.cost1 = NULL; }
        d.cond = NULL;
        if (j.cond_store >= 0)
        {
            d.cond = (const int*)mvStorePtr(c, j.b_slot, j.cond_store) + 2 * g.ncu;
            st = batchWaitSearches(c, b, c->mvWriter[(size_t)j.b_slot * nmv + j.cond_store]);
            if (st) return st;
        }
        d.intraCost = slotPtr<int>(c, j.b_slot, L.intraCost);
        d.invQ = c->cfg.need_aq ? slotInvQ(c, j.b_slot) : NULL;
        d.lowresCosts = (unsigned short*)cs;
        d.rowSatds = (int*)(cs + L.costRowOff);
        d.result = (CostResultDev*)(cs + L.costResOff);
    }
    n = nMine;
    if (!n) return implicit ? endBatch(c) : X265CU_OK;
    c->costEnq += n;
    /* B estimates that share the source frame and the list-1 search form one group (cost_group_kernel) */
    std::vector<CostGroupDev<P> > groups;
    {
        std::map<std::pair<const void*, const void*>, int> open;       /* (fenc planes, list-1 MVs) -> group being filled */
        for (int i = nP; i < n; i++)
        {
            const CostJobDev<P>& d = dev[i];
            const std::pair<const void*, const void*> key((const void*)d.fenc0, (const void*)d.mv1);
            std::map<std::pair<const void*, const void*>, int>::iterator it = open.find(key);
            int gi;
            if (it == open.end() || groups[it->second].n == LA_COST_GROUP_MAX)
            {
                gi = (int)groups.size();
                CostGroupDev<P> G;
                memset(&G, 0, sizeof(G));
                G.fenc0 = d.fenc0; G.ref1 = d.ref1; G.mv1 = d.mv1; G.cost1 = d.cost1; G.intraCost = d.intraCost; G.invQ = d.invQ;
                groups.push_back(G);
                open[key] = gi;
            }
            else
                gi = it->second;
            CostGroupDev<P>& G = groups[gi];
            typename CostGroupDev<P>::Member& M = G.m[G.n++];
            M.ref0 = d.ref0; M.mv0 = d.mv0; M.cost0 = d.cost0; M.lowresCosts = d.lowresCosts; M.rowSatds = d.rowSatds;
            M.result = d.result; M.cond = d.cond;
        }
    }
    char *hgr = NULL, *dgr = NULL;
    if (!groups.empty())
    {
        st = batchStage(c, b, groups.size() * sizeof(CostGroupDev<P>), &hgr, &dgr);
        if (st) return st;
        memcpy(hgr, &groups[0], groups.size() * sizeof(CostGroupDev<P>));
    }
    {
        void* hAlias = hostAlias(hst);
        void* gAlias = hgr ? hostAlias(hgr) : NULL;
        if (!hAlias || (hgr && !gAlias)) { snprintf(c->err, sizeof(c->err), "job staging memory is not device-visible"); return X265CU_ERR_CUDA; }
        st = stageLaunch(c, b->stream, dst, hAlias, alignUp(n * sizeof(CostJobDev<P>), 16), dgr, gAlias,
                         alignUp(groups.size() * sizeof(CostGroupDev<P>), 16), NULL, 0);
        if (st) return st;
    }
    {
        Prof pr(c, X265CU_K_COST, 1 + (nP > 0) + (n > nP), b->stream);
        for (int base = 0; base < n; base += 65535)
            cost_clear_kernel<P><<<(n - base) < 65535 ? (n - base) : 65535, 256, 0, b->stream>>>(g, (const CostJobDev<P>*)dst + base, c->d_executed + 1);
        for (int base = 0; base < nP; base += 65535)
        {
            dim3 gg((g.ncu + 127) / 128, (unsigned)((nP - base) < 65535 ? (nP - base) : 65535));
            cost_p_kernel<P><<<gg, 128, 0, b->stream>>>(g, (const CostJobDev<P>*)dst + base);
        }
        const int ng = (int)groups.size();
        for (int base = 0; base < ng; base += 65535)
        {
            const unsigned ny = (unsigned)((ng - base) < 65535 ? (ng - base) : 65535);
            if (c->costLanes == 8) cost_group_kernel<P><<<dim3((g.ncu + 15) / 16, ny), 128, 0, b->stream>>>(g, (const CostGroupDev<P>*)dgr + base);
            else cost_group_kernel4<P><<<dim3((g.ncu + 31) / 32, ny), 128, 0, b->stream>>>(g, (const CostGroupDev<P>*)dgr + base);
        }
    }
    CK(cudaGetLastError());
    if (implicit) return endBatch(c);
    return X265CU_OK;
}

template <typename P>
int weightCostT(x265cu_ctx* c, const x265cu_wcost_job* jobs, int n, uint32_t* costs)
{
    HostTimer ht(HT_WEIGHT);
    const Geom& g = c->g;
    const SlotLayout& L = c->lay;
    int st = ensureDev(c, &c->d_results, &c->resultsCap, n * sizeof(unsigned));
    if (st) return st;
    CK(cudaMemsetAsync(c->d_results, 0, n * sizeof(unsigned), c->stream));
    st = ensureScratch(c, c->mainScratch, c->stream, 1);
    if (st) return st;
    for (int i = 0; i < n; i++)
    {
        const x265cu_wcost_job& j = jobs[i];
        if (!slotOk(c, j.fenc_slot) || !slotOk(c, j.ref_slot)) return X265CU_ERR_BAD_ARG;
        st = mainWaitPre(c, j.fenc_slot);
        if (!st) st = mainWaitPre(c, j.ref_slot);
        if (st) return st;
        const P* ref = slotPtr<P>(c, j.ref_slot, L.planes);
        if (j.weighted)
        {
            st = weightPlanes<P>(c, c->stream, ref, (P*)c->mainScratch[0], 1, j.w_scale, j.w_denom, j.w_offset);
            if (st) return st;
            ref = (const P*)c->mainScratch[0];
        }
        Prof pr(c, X265CU_K_WEIGHT, 1);
        weight_cost_kernel<P><<<(g.ncu + 15) / 16, 128, 0, c->stream>>>(g, slotPtr<P>(c, j.fenc_slot, L.planes),
                                                                        ref, slotPtr<int>(c, j.fenc_slot, L.intraCost),
                                                                        (unsigned*)c->d_results + i);
    }
    CK(cudaGetLastError());
    st = ensureMapped(c, n * sizeof(unsigned));
    if (st) return st;
    publish_kernel<<<1, 128, 0, c->stream>>>((const unsigned*)c->d_results, (unsigned*)c->d_mapped, n);
    c->counters.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    memcpy(costs, c->h_mapped, n * sizeof(unsigned));
    c->counters.d2h_bytes += n * sizeof(unsigned);
    return X265CU_OK;
}

template <typename P>
int blockMetricsT(x265cu_ctx* c, const void* a, const void* b, int n, int32_t* sad, int32_t* satd)
{
    P *da = NULL, *db = NULL; int *ds = NULL, *dt = NULL;
    const size_t bytes = (size_t)n * 64 * sizeof(P);
    CK(cudaMalloc((void**)&da, bytes + 64)); CK(cudaMalloc((void**)&db, bytes + 64));
    CK(cudaMalloc((void**)&ds, n * sizeof(int))); CK(cudaMalloc((void**)&dt, n * sizeof(int)));
    CK(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, c->stream));
    block_metrics_kernel<P><<<(n + 15) / 16, 128, 0, c->stream>>>(da, db, n, ds, dt);
    c->counters.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(sad, ds, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(satd, dt, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(da); cudaFree(db); cudaFree(ds); cudaFree(dt);
    return X265CU_OK;
}

template <typename P>
int mcMetricsT(x265cu_ctx* c, int fencSlot, int refSlot, const int32_t* cuIdx, const int32_t* mvs, int n, int32_t* sad, int32_t* satd)
{
    int *dc = NULL, *dm = NULL, *ds = NULL, *dt = NULL;
    CK(cudaMalloc((void**)&dc, n * sizeof(int))); CK(cudaMalloc((void**)&dm, 2 * n * sizeof(int)));
    CK(cudaMalloc((void**)&ds, n * sizeof(int))); CK(cudaMalloc((void**)&dt, n * sizeof(int)));
    CK(cudaMemcpyAsync(dc, cuIdx, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dm, mvs, 2 * n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    mc_metrics_kernel<P><<<(n + 15) / 16, 128, 0, c->stream>>>(c->g, slotPtr<P>(c, fencSlot, c->lay.planes), slotPtr<P>(c, refSlot, c->lay.planes),
                                                               dc, dm, n, ds, dt);
    c->counters.kernel_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(sad, ds, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(satd, dt, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(dc); cudaFree(dm); cudaFree(ds); cudaFree(dt);
    return X265CU_OK;
}

#define DISPATCH(fn, ...) (c->bpp == 1 ? fn<uint8_t>(__VA_ARGS__) : c->cfg.depth > 10 ? fn<px12>(__VA_ARGS__) : fn<uint16_t>(__VA_ARGS__))

} // namespace

extern "C" {

int x265cu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* x265cu_strerror(int s)
{
    switch (s)
    {
    case X265CU_OK: return "ok";
    case X265CU_ERR_NO_DEVICE: return "no usable CUDA device (the GPU lookahead has no CPU fallback)";
    case X265CU_ERR_BAD_ARG: return "bad argument";
    case X265CU_ERR_NO_MEMORY: return "out of memory";
    case X265CU_ERR_CUDA: return "CUDA error";
    case X265CU_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
    }
}

const char* x265cu_last_error(const x265cu_ctx* c) { return c ? c->err : ""; }

static int createImpl(const x265cu_config* cfg, x265cu_ctx** out);

int x265cu_create(const x265cu_config* cfg, x265cu_ctx** out)
{
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    const int rc = createImpl(cfg, out);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

static int createImpl(const x265cu_config* cfg, x265cu_ctx** out)
{
    if (!cfg || !out) return X265CU_ERR_BAD_ARG;
    *out = NULL;
    if (cfg->qg_size != 8 && cfg->qg_size != 16 && cfg->qg_size != 32 && cfg->qg_size != 64) return X265CU_ERR_BAD_ARG;
    if (cfg->depth != 8 && cfg->depth != 10 && cfg->depth != 12) return X265CU_ERR_UNSUPPORTED;     /* 12-bit: la_device.cuh px12 */
    if (cfg->hist_stats && (cfg->depth != 8 || cfg->width < 64 || cfg->height < 64))
        return X265CU_ERR_UNSUPPORTED;      /* the reference indexes 256 histogram bins with the sample value */
    if (cfg->fade_stats)
    {
        /* --fades: the second acEnergyCu pass (slicetype.cpp:697-712) walks the picture ROUNDED to 16 when weightp is on
         * (:683-684).  Supported where that is the AQ grid itself: 16x16 AQ blocks, and either no weightp or a size whose
         * remainder modulo 16 is 0 or >= 8 (1080, 2160, 720, 360, ...); anything else would need per-block sums of a sub-grid */
        const bool wp = cfg->need_wp_stats != 0;
        const int rw = cfg->width & 15, rh = cfg->height & 15;
        if (!cfg->need_aq || cfg->qg_size == 8 || (wp && ((rw > 0 && rw < 8) || (rh > 0 && rh < 8))) ||
            (cfg->height + 15) / 16 + 1 > LA_FADE_MAX_ROWS)
            return X265CU_ERR_UNSUPPORTED;
    }
    if (cfg->hme)
    {
        /* dia / hex / umh / star / full, not sea (la_me_generic.cuh).  No cooperative slices: the reference's two levels race there (its level-1
         * slices read level-0 vectors other workers may not have written).  CTU >= 32: the level-0 margins (half the lowres
         * ones) must be whole tiles and cover what a search can reach */
        for (int i = 0; i < 2; i++)
            if (cfg->hme_search[i] < 0 || cfg->hme_search[i] > 5 || cfg->hme_search[i] == 4 || cfg->hme_range[i] < 4 || cfg->hme_range[i] > 256) return X265CU_ERR_UNSUPPORTED;
        if (cfg->rows_per_slice > 0 || cfg->max_cu_size < 32 || cfg->width < 64 || cfg->height < 64) return X265CU_ERR_UNSUPPORTED;
    }
    if (cfg->aq_mode < 0 || cfg->aq_mode > 5 || (cfg->aq_mode > 3 && cfg->fade_stats)) return X265CU_ERR_UNSUPPORTED;
    if (cfg->width < 16 || cfg->height < 16 || cfg->bframes < 0 || cfg->bframes > 16 || cfg->max_slots < 1 || !cfg->mvcost)
        return X265CU_ERR_BAD_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device >= ndev) return X265CU_ERR_NO_DEVICE;
    if (cudaSetDevice(cfg->device) != cudaSuccess) return X265CU_ERR_NO_DEVICE;
    x265cu_ctx* c = new (std::nothrow) x265cu_ctx();
    if (!c) return X265CU_ERR_NO_MEMORY;
    c->cfg = *cfg; c->err[0] = 0; c->d_mvcost = NULL; c->d_executed = NULL;
    c->d_histAcc = NULL; c->h_hist = NULL; c->d_hist = NULL;
    c->d_results = NULL; c->resultsCap = 0; c->h_slotStats = NULL; c->d_slotStats = NULL; c->h_mapped = NULL; c->d_mapped = NULL; c->mappedCap = 0;
    c->h_ctJobs = NULL; c->d_ctJobs = NULL; c->ctRingPos = 0; c->ctPending = 0;
    c->searchWorkers = getenv("X265CU_SEARCH_WORKERS") ? atoi(getenv("X265CU_SEARCH_WORKERS")) : 0;
    c->searchSmem = getenv("X265CU_SEARCH_SMEM") ? atoi(getenv("X265CU_SEARCH_SMEM")) : 0;
    c->searchLanes = getenv("X265CU_SEARCH_LANES") && atoi(getenv("X265CU_SEARCH_LANES")) == 4 ? 4 : 8;
    c->searchOneShot = getenv("X265CU_SEARCH_ONESHOT") ? atoi(getenv("X265CU_SEARCH_ONESHOT")) : 0;
    c->bigSearch[0] = c->bigSearch[1] = -1;
    c->costLanes = getenv("X265CU_COST_LANES") && atoi(getenv("X265CU_COST_LANES")) == 8 ? 8 : 4;
    c->numSMs = 148;
    cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, cfg->device);
    c->profile = false; c->evUsed = 0; c->nextBatch = 0; c->cur = NULL; c->searchEnq = c->costEnq = 0;
    c->rank = 0; c->nranks = 1; c->exchange = NULL; c->exchangeUser = NULL;
    c->mirrorStream = NULL; c->gatherStream = NULL; c->mirrorMark = NULL; c->nextMirror = 0; c->d_recalc = c->h_recalc = NULL; c->recalcStride = 0;
    for (int i = 0; i < X265CU_MIRROR_RING; i++) { c->mirror[i].done = NULL; c->mirror[i].scratch = NULL; c->mirror[i].cap = 0; c->mirror[i].ticket = -1; }
    c->stream = c->copyStream = NULL; c->preSeq = 0; for (int i = 0; i < LA_NUM_PRE; i++) c->preStreams[i] = NULL; c->profBase = c->tm0 = c->tm1 = c->mainMark = NULL;
    for (int i = 0; i < LA_NUM_LANES; i++) c->lanes[i] = NULL;
    for (int i = 0; i < LA_NUM_BATCHES; i++)
    {
        Batch& b = c->batches[i];
        b.id = -1; b.stream = NULL; b.searchDone = b.done = NULL; b.open = false; for (int k = 0; k < LA_NUM_PRE; k++) b.begun[k] = NULL;
        b.h_stage = b.d_stage = NULL; b.stageCap = b.stageUsed = 0; b.d_sync = NULL; b.syncCap = b.syncUsed = 0;
        for (int r = 0; r < LA_MAX_RANKS; r++) { b.xbuf[r] = NULL; b.xcap[r] = 0; }
    }
    memset(&c->counters, 0, sizeof(c->counters)); memset(c->profMs, 0, sizeof(c->profMs)); memset(c->profN, 0, sizeof(c->profN));
    memset(c->profBusy, 0, sizeof(c->profBusy));
    c->bpp = cfg->depth > 8 ? 2 : 1;

    /* geometry: Lowres::create (lowres.cpp:72-97) */
    Geom& g = c->g;
    g.picW = cfg->width; g.picH = cfg->height; g.cW = (cfg->width + 1) / 2; g.cH = (cfg->height + 1) / 2;
    g.srcPitch = (g.picW + 15) / 16 * 16; g.srcPitchC = (g.cW + 15) / 16 * 16;
    const int lw = cfg->width / 2, lh = cfg->height / 2;
    g.mx = cfg->max_cu_size + 32; g.my = cfg->max_cu_size + 16;
    g.stride = lw + 2 * g.mx;
    if (g.stride & 31) g.stride += 32 - (g.stride & 31);
    g.bw = (lw + 7) >> 3; g.bh = (lh + 7) >> 3; g.ncu = g.bw * g.bh;
    g.w = g.bw * 8; g.h = g.bh * 8;
    g.planeLines = g.h + 2 * g.my;
    g.planeSize = (long long)g.stride * g.planeLines;
    g.padOffset = (long long)g.stride * g.my + g.mx;
    g.lambda = cfg->lambda; g.depth = cfg->depth; g.nb = cfg->bframes + 2;
    g.tpr = g.stride / 8;
    g.rowsPerSlice = cfg->rows_per_slice > 0 ? cfg->rows_per_slice : 0;
    /* calcAdaptiveQuantFrame's block grid (slicetype.cpp:459-472, lowres.cpp:86-89) */
    g.aqBlock = cfg->qg_size == 8 ? 8 : 16;
    g.aqW = (g.picW + g.aqBlock - 1) / g.aqBlock; g.aqH = (g.picH + g.aqBlock - 1) / g.aqBlock;
    g.ncuFull = cfg->qg_size == 8 ? 4 * g.ncu : g.ncu;
    /* --hme level 0 (lowres.cpp:165-183, 378-388; slicetype.cpp:998-999): planes of w / 2 x h / 2 samples with half the
     * margins; the buffer is a little larger than the reference's (whole tiles, and room for the clamped loads) */
    Geom& g4 = c->g4;
    g4 = g;
    g4.w = g.w / 2; g4.h = g.h / 2;
    g4.bw = ((cfg->width / 4) + 7) >> 3; g4.bh = ((cfg->height / 4) + 7) >> 3; g4.ncu = g4.bw * g4.bh;
    g4.mx = g.mx / 2; g4.my = g.my / 2;
    g4.stride = (std::max(g4.w, 8 * g4.bw) + 2 * g4.mx + 16 + 7) & ~7;
    g4.planeLines = (std::max(g4.h, 8 * g4.bh) + 2 * g4.my + 16 + 7) & ~7;
    g4.planeSize = (long long)g4.stride * g4.planeLines;
    g4.padOffset = (long long)g4.stride * g4.my + g4.mx;
    g4.tpr = g4.stride / 8;
    g4.rowsPerSlice = 0;
    x265cu_geometry& G = c->geom;
    G.low_width = g.w; G.low_height = g.h; G.bw = g.bw; G.bh = g.bh; G.ncu = g.ncu; G.stride = g.stride;
    G.plane_lines = g.planeLines; G.margin_x = g.mx; G.margin_y = g.my; G.nb = g.nb;
    G.n_mv_stores = (cfg->mv_store_kinds > 0 ? cfg->mv_store_kinds : 3) * g.nb;
    G.n_cost_stores = (cfg->cost_variants > 0 ? cfg->cost_variants : 2) * g.nb * g.nb;
    G.ncu_full = g.ncuFull;

    SlotLayout& L = c->lay;
    size_t o = 0;
#define SECTION(name, bytes) L.name = o; o = alignUp(o + (bytes), 256)
    SECTION(srcY, (size_t)g.srcPitch * g.picH * c->bpp + 1024);     /* + slack: K1's row copies may run past the last row's end */
    SECTION(srcU, (size_t)g.srcPitchC * g.cH * c->bpp + 512);
    SECTION(srcV, (size_t)g.srcPitchC * g.cH * c->bpp + 512);
    SECTION(planes, (size_t)(4 * g.planeSize) * c->bpp + 256);
    SECTION(intraCost, (size_t)g.ncu * 4);
    SECTION(intraMode, (size_t)g.ncu);
    /* with a ragged picture the running AQ index can visit a few more blocks than ncuFull holds: slack of one row */
    const size_t nAq = (size_t)std::max(g.ncuFull, g.aqW * g.aqH) + 2 * g.bw + 2;
    SECTION(invQ, nAq * 4);
    SECTION(invQ8, (size_t)g.ncu * 4);
    SECTION(qpAq, nAq * 8);
    SECTION(qpCuTree, nAq * 8);
    SECTION(propagate, (size_t)g.ncu * 4);
    SECTION(energy, nAq * 4);
    SECTION(edge, cfg->aq_mode > 3 ? nAq * 4 : 0);
    SECTION(aqSums, 16);
    SECTION(lowresCosts00, (size_t)g.ncu * 2);
    L.rowSatds00 = o; o += alignUp((size_t)g.bh * 4, 16);
    L.stats = o; o = alignUp(o + sizeof(FrameStatsDev), 256);
    L.mvStoreStride = alignUp((size_t)g.ncu * 8 + 16, 256);     /* packed MVs, costs, skip flag */
    SECTION(mvStores, L.mvStoreStride * G.n_mv_stores);
    L.costRowOff = alignUp((size_t)g.ncu * 2, 16);
    L.costResOff = L.costRowOff + alignUp((size_t)g.bh * 4, 16);
    L.costStoreStride = alignUp(L.costResOff + sizeof(CostResultDev), 256);
    SECTION(costStores, L.costStoreStride * G.n_cost_stores);
    L.mvStore4Stride = alignUp((size_t)g4.ncu * 8, 256);        /* level-0 packed MVs + costs of one (list, distance) */
    if (cfg->hme)
    {
        SECTION(planes4, (size_t)(4 * g4.planeSize) * c->bpp + 256);
        SECTION(mvStores4, L.mvStore4Stride * G.n_mv_stores);
    }
    else { L.planes4 = L.mvStores4 = 0; }
#undef SECTION
    L.total = o;

    int rc = X265CU_OK;
    /* The worker lanes run at the lowest priority: their search warps live for milliseconds and would otherwise
     * leave the short pre-lookahead / cuTree kernels (which the host waits for) queueing for a free SM slot */
    int prLeast = 0, prGreatest = 0;
    cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest);
    c->greenSmall = c->greenLarge = NULL; c->greenSmallSMs = c->greenLargeSMs = 0;
    {
        const char* e = getenv("X265CU_GREEN");
        const int want = e ? atoi(e) : 16;      /* SMs of the small partition; 0 = no partitioning */
        if (want > 0 && !makeGreenStreams(c, want, prGreatest, prLeast))
        {
            c->greenSmall = c->greenLarge = NULL;
            c->stream = c->mirrorStream = c->gatherStream = NULL;
            for (int i = 0; i < LA_NUM_LANES; i++) c->lanes[i] = NULL;
        }
    }
    if (!c->stream && cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prGreatest) != cudaSuccess) { delete c; return X265CU_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return X265CU_ERR_CUDA; }
    for (int i = 0; i < LA_NUM_PRE; i++)
        if (cudaStreamCreateWithPriority(&c->preStreams[i], cudaStreamNonBlocking, prGreatest) != cudaSuccess) rc = X265CU_ERR_CUDA;
    if (cudaEventCreateWithFlags(&c->mainMark, cudaEventDisableTiming) != cudaSuccess) rc = X265CU_ERR_CUDA;
    if ((!c->gatherStream && cudaStreamCreateWithPriority(&c->gatherStream, cudaStreamNonBlocking, prGreatest) != cudaSuccess) ||
        (!c->mirrorStream && cudaStreamCreateWithPriority(&c->mirrorStream, cudaStreamNonBlocking, prGreatest) != cudaSuccess) ||
        cudaEventCreateWithFlags(&c->mirrorMark, cudaEventDisableTiming) != cudaSuccess) rc = X265CU_ERR_CUDA;
    for (int i = 0; !rc && i < X265CU_MIRROR_RING; i++)
        if (cudaEventCreateWithFlags(&c->mirror[i].done, cudaEventDisableTiming) != cudaSuccess) rc = X265CU_ERR_CUDA;
    for (int i = 0; !rc && i < LA_NUM_LANES; i++)
        if (!c->lanes[i] && cudaStreamCreateWithPriority(&c->lanes[i], cudaStreamNonBlocking, prLeast) != cudaSuccess) rc = X265CU_ERR_CUDA;
    for (int i = 0; !rc && i < LA_NUM_BATCHES; i++)
    {
        Batch& b = c->batches[i];
        for (int k = 0; k < LA_NUM_PRE; k++)
            if (cudaEventCreateWithFlags(&b.begun[k], cudaEventDisableTiming) != cudaSuccess) rc = X265CU_ERR_CUDA;
        if (
            cudaEventCreateWithFlags(&b.searchDone, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming) != cudaSuccess) rc = X265CU_ERR_CUDA;
    }
    cudaEventCreate(&c->tm0); cudaEventCreate(&c->tm1); cudaEventCreate(&c->profBase);
    if (!rc && cudaMalloc((void**)&c->d_executed, 2 * sizeof(unsigned long long)) != cudaSuccess) rc = X265CU_ERR_NO_MEMORY;
    if (!rc && cudaMemset(c->d_executed, 0, 2 * sizeof(unsigned long long)) != cudaSuccess) rc = X265CU_ERR_CUDA;
    const size_t tabBytes = (2 * (size_t)cfg->mvcost_half + 1) * sizeof(unsigned short);
    if (cudaMalloc((void**)&c->d_mvcost, tabBytes) != cudaSuccess) rc = X265CU_ERR_NO_MEMORY;
    if (!rc && cudaMemcpy(c->d_mvcost, cfg->mvcost, tabBytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = X265CU_ERR_CUDA;
    for (int i = 0; !rc && i < cfg->max_slots; i++)
    {
        char* p = NULL;
        if (cudaMalloc((void**)&p, L.total) != cudaSuccess) { rc = X265CU_ERR_NO_MEMORY; break; }
        c->slots.push_back(p);
        cudaEvent_t e0, e1;
        cudaEventCreateWithFlags(&e0, cudaEventDisableTiming); cudaEventCreateWithFlags(&e1, cudaEventDisableTiming);
        c->slotCopied.push_back(e0); c->slotConsumed.push_back(e1);
        c->slotUsers.push_back(std::vector<long long>());
        c->slotOwner.push_back(0);
        c->slotMainTouched.push_back(0);
        cudaEvent_t e2, e3;
        cudaEventCreateWithFlags(&e2, cudaEventDisableTiming); cudaEventCreateWithFlags(&e3, cudaEventDisableTiming);
        c->slotMirrored.push_back(e2); c->slotMirrorTouched.push_back(0);
        c->slotRecalc.push_back(e3); c->slotRecalcStore.push_back(-1);
        cudaEvent_t e4;
        cudaEventCreateWithFlags(&e4, cudaEventDisableTiming);
        c->slotMainWrote.push_back(e4); c->slotMainWroteSet.push_back(0);
        /* planes must start zeroed: columns past the right margin are never written (K1) */
        if (cudaMemsetAsync(p, 0, L.total, c->stream) != cudaSuccess) rc = X265CU_ERR_CUDA;
    }
    /* everything a steady-state batch needs exists before the first frame arrives: cudaMalloc / cudaMallocHost in
     * the middle of a run stall the host for milliseconds while kernels are in flight */
    for (int i = 0; !rc && i < LA_NUM_BATCHES; i++)
    {
        Batch& b = c->batches[i];
        /* room for a batch of ~96 frames (the first decision's): per frame 3(B+1) search jobs, (B+1)(B+4) cost jobs
         * with their groups, and the tickets / progress counters of its searches */
        const size_t B1 = (size_t)cfg->bframes + 1;
        const size_t nstripsEst = (size_t)(g.bh + LA_STRIP_ROWS - 1) / LA_STRIP_ROWS;
        const size_t perFrameStage = 3 * B1 * 64 + B1 * (B1 + 3) * 104 + 2 * B1 * 1024;
        const size_t perFrameSync = 3 * B1 * nstripsEst * sizeof(int) + 64;
        b.stageCap = alignUp(std::max((size_t)256 << 10, perFrameStage * 96), 4096);
        b.syncCap = alignUp(std::max((size_t)64 << 10, perFrameSync * 96), 4096);
        if (cudaHostAlloc((void**)&b.h_stage, b.stageCap, cudaHostAllocMapped) != cudaSuccess || cudaMalloc((void**)&b.d_stage, b.stageCap) != cudaSuccess ||
            cudaMalloc((void**)&b.d_sync, b.syncCap) != cudaSuccess) rc = X265CU_ERR_NO_MEMORY;
    }
    if (!rc && cfg->need_wp_stats && ensureScratch(c, c->mainScratch, c->stream, 1) != X265CU_OK) rc = X265CU_ERR_NO_MEMORY;
    if (!rc && cfg->hist_stats)
    {
        if (cudaMalloc((void**)&c->d_histAcc, (size_t)cfg->max_slots * sizeof(HistAcc)) != cudaSuccess ||
            cudaHostAlloc((void**)&c->h_hist, (size_t)cfg->max_slots * sizeof(HistStatsDev), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void**)&c->d_hist, c->h_hist, 0) != cudaSuccess) rc = X265CU_ERR_NO_MEMORY;
    }
    if (!rc)
    {
        c->recalcStride = alignUp(8 + (size_t)g.bh * 4, 256);
        if (cudaMalloc((void**)&c->d_recalc, c->recalcStride * cfg->max_slots) != cudaSuccess ||
            cudaHostAlloc((void**)&c->h_recalc, c->recalcStride * cfg->max_slots, cudaHostAllocMapped) != cudaSuccess) rc = X265CU_ERR_NO_MEMORY;
    }
    if (!rc)
    {
        c->resultsCap = (size_t)1 << 20;
        if (cudaMalloc((void**)&c->d_results, c->resultsCap) != cudaSuccess ||
            cudaHostAlloc((void**)&c->h_slotStats, cfg->max_slots * sizeof(FrameStatsDev), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void**)&c->d_slotStats, c->h_slotStats, 0) != cudaSuccess ||
            cudaHostAlloc((void**)&c->h_ctJobs, LA_CT_RING * sizeof(CutreeJobDev), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void**)&c->d_ctJobs, c->h_ctJobs, 0) != cudaSuccess ||
            ensureMapped(c, (size_t)256 << 10) != X265CU_OK) rc = X265CU_ERR_NO_MEMORY;
    }
    if (!rc)
    {
        c->mvWriter.assign(c->slots.size() * (size_t)G.n_mv_stores, -1LL);
        c->costWriter.assign(c->slots.size() * (size_t)G.n_cost_stores, -1LL);
        cudaEventRecord(c->profBase, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = X265CU_ERR_CUDA;
    }
    if (rc) { x265cu_destroy(c); return rc; }
    c->cfg.mvcost = NULL;
    *out = c;
    return X265CU_OK;
}

void x265cu_destroy(x265cu_ctx* c)
{
    if (c && getenv("X265CU_HOST_TIMING"))
    {
        fprintf(stderr, "x265cu host timing (ms, calls):");
        for (int i = 0; i < HT_COUNT; i++) { fprintf(stderr, " %s %.1f/%llu", g_htNames[i], g_ht[i] * 1e3, g_htN[i]); g_ht[i] = 0; g_htN[i] = 0; }
        fprintf(stderr, "\n");
    }
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!c) return;
    syncAll(c);
    for (size_t i = 0; i < c->slots.size(); i++) cudaFree(c->slots[i]);
    for (size_t i = 0; i < c->slotCopied.size(); i++) { cudaEventDestroy(c->slotCopied[i]); cudaEventDestroy(c->slotConsumed[i]); }
    for (size_t i = 0; i < c->slotMirrored.size(); i++) { cudaEventDestroy(c->slotMirrored[i]); cudaEventDestroy(c->slotRecalc[i]); }
    for (size_t i = 0; i < c->slotMainWrote.size(); i++) cudaEventDestroy(c->slotMainWrote[i]);
    for (int i = 0; i < X265CU_MIRROR_RING; i++) { if (c->mirror[i].done) cudaEventDestroy(c->mirror[i].done); cudaFree(c->mirror[i].scratch); }
    if (c->mirrorStream) cudaStreamDestroy(c->mirrorStream);
    if (c->gatherStream) cudaStreamDestroy(c->gatherStream);
    if (c->mirrorMark) cudaEventDestroy(c->mirrorMark);
    cudaFree(c->d_recalc); if (c->h_recalc) cudaFreeHost(c->h_recalc);
    {
        std::lock_guard<std::mutex> lock(g_evMutex);
        std::vector<std::pair<cudaEvent_t, cudaEvent_t> >& fr = g_evFree[c->cfg.device];
        for (size_t i = 0; i < c->evPool.size(); i++)
        {
            if (fr.size() < 16384) fr.push_back(std::make_pair(c->evPool[i].a, c->evPool[i].b));
            else { cudaEventDestroy(c->evPool[i].a); cudaEventDestroy(c->evPool[i].b); }
        }
    }
    for (size_t i = 0; i < c->mainScratch.size(); i++) cudaFree(c->mainScratch[i]);
    for (int i = 0; i < LA_NUM_BATCHES; i++)
    {
        Batch& b = c->batches[i];
        for (size_t k = 0; k < b.weightScratch.size(); k++) cudaFree(b.weightScratch[k]);
        for (size_t k = 0; k < b.retiredDev.size(); k++) cudaFree(b.retiredDev[k]);
        for (size_t k = 0; k < b.retiredHost.size(); k++) cudaFreeHost(b.retiredHost[k]);
        cudaFree(b.d_stage); cudaFree(b.d_sync);
        for (int r = 0; r < LA_MAX_RANKS; r++) cudaFree(b.xbuf[r]);
        if (b.h_stage) cudaFreeHost(b.h_stage);
        for (int k = 0; k < LA_NUM_PRE; k++) if (b.begun[k]) cudaEventDestroy(b.begun[k]);
        if (b.searchDone) cudaEventDestroy(b.searchDone);
        if (b.done) cudaEventDestroy(b.done);
    }
    for (size_t i = 0; i < c->xpool.size(); i++) cudaFree(c->xpool[i].first);
    cudaFree(c->d_mvcost); cudaFree(c->d_executed); cudaFree(c->d_results);
    if (c->h_slotStats) cudaFreeHost(c->h_slotStats);
    cudaFree(c->d_histAcc); if (c->h_hist) cudaFreeHost(c->h_hist);
    if (c->h_mapped) cudaFreeHost(c->h_mapped);
    if (c->h_ctJobs) cudaFreeHost(c->h_ctJobs);
    if (c->tm0) cudaEventDestroy(c->tm0);
    if (c->tm1) cudaEventDestroy(c->tm1);
    if (c->profBase) cudaEventDestroy(c->profBase);
    for (int i = 0; i < LA_NUM_LANES; i++) if (c->lanes[i]) cudaStreamDestroy(c->lanes[i]);
    for (int i = 0; i < LA_NUM_PRE; i++) if (c->preStreams[i]) cudaStreamDestroy(c->preStreams[i]);
    if (c->mainMark) cudaEventDestroy(c->mainMark);
    if (c->copyStream) cudaStreamDestroy(c->copyStream);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->greenSmall || c->greenLarge)
    {
        typedef CUresult (*GreenDestroy)(CUgreenCtx);
        void* f = NULL;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuGreenCtxDestroy", &f, cudaEnableDefault, &q) == cudaSuccess && f)
        {
            if (c->greenSmall) ((GreenDestroy)f)((CUgreenCtx)c->greenSmall);
            if (c->greenLarge) ((GreenDestroy)f)((CUgreenCtx)c->greenLarge);
        }
    }
    delete c;
}

int x265cu_sm_partition(const x265cu_ctx* c, int32_t* small_sms, int32_t* large_sms)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    if (small_sms) *small_sms = c->greenSmallSMs;
    if (large_sms) *large_sms = c->greenLargeSMs;
    return X265CU_OK;
}

int x265cu_get_geometry(const x265cu_ctx* c, x265cu_geometry* out) { if (!c || !out) return X265CU_ERR_BAD_ARG; *out = c->geom; return X265CU_OK; }

int x265cu_pin_host(x265cu_ctx* c, void* ptr, uint64_t bytes)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    c->mappedCache.clear();
    return X265CU_OK;
}
int x265cu_unpin_host(x265cu_ctx* c, void* ptr)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    CK(cudaHostUnregister(ptr));
    c->mappedCache.clear();
    return X265CU_OK;
}

/* the main stream waits for every batch in flight (so an event recorded on it afterwards covers all the work) */
static int mainJoinBatches(x265cu_ctx* c)
{
    int st = endBatch(c);
    if (st) return st;
    for (int i = 0; i < LA_NUM_BATCHES; i++)
        if (c->batches[i].id >= 0) CK(cudaStreamWaitEvent(c->stream, c->batches[i].done, 0));
    return X265CU_OK;
}

int x265cu_sync(x265cu_ctx* c)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    int st = endBatch(c);
    if (st) return st;
    CK(cudaStreamSynchronize(c->copyStream)); for (int i = 0; i < LA_NUM_PRE; i++) CK(cudaStreamSynchronize(c->preStreams[i])); CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < LA_NUM_LANES; i++) CK(cudaStreamSynchronize(c->lanes[i]));
    CK(cudaStreamSynchronize(c->mirrorStream));
    return X265CU_OK;
}

int x265cu_batch_begin(x265cu_ctx* c, int64_t* batch_id)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!c) return X265CU_ERR_BAD_ARG;
    int st = beginBatch(c);
    if (!st && batch_id) *batch_id = c->cur->id;
    return st;
}

int x265cu_batches_in_flight(x265cu_ctx* c)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    /* newest first, one lane's worth: the answer the caller acts on is "fewer than two", so counting stops there */
    int n = 0;
    for (long long id = c->nextBatch - 1; id >= 0 && id >= c->nextBatch - LA_NUM_LANES && n < 2; id--)
    {
        const Batch* b = batchOf(c, id);
        if (b && (b->open || cudaEventQuery(b->done) == cudaErrorNotReady)) n++;
    }
    cudaGetLastError();
    return n;
}

int x265cu_batch_end(x265cu_ctx* c)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    return endBatch(c);
}

int x265cu_shard_config(x265cu_ctx* c, int32_t rank, int32_t nranks, x265cu_exchange_fn fn, void* user)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!c || nranks < 1 || nranks > LA_MAX_RANKS || rank < 0 || rank >= nranks || (nranks > 1 && !fn)) return X265CU_ERR_BAD_ARG;
    int st = x265cu_sync(c);
    if (st) return st;
    c->rank = rank; c->nranks = nranks; c->exchange = fn; c->exchangeUser = user;
    return X265CU_OK;
}

int x265cu_slot_owner(x265cu_ctx* c, int32_t slot, int32_t owner)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    if (!c || !slotOk(c, slot) || owner < 0 || owner >= c->nranks) return X265CU_ERR_BAD_ARG;
    c->slotOwner[slot] = owner;
    return X265CU_OK;
}

/* device-side stopwatch (bench.py times its steps with it): start on the main stream; stop after everything
 * enqueued on any stream of the context */
int x265cu_timer_start(x265cu_ctx* c)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    CK(cudaEventRecord(c->tm0, c->stream));
    return X265CU_OK;
}
int x265cu_timer_stop(x265cu_ctx* c, double* ms)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    int st = mainJoinBatches(c);
    if (st) return st;
    CK(cudaStreamSynchronize(c->copyStream));
    for (int i = 0; i < LA_NUM_PRE; i++) CK(cudaStreamSynchronize(c->preStreams[i]));
    CK(cudaStreamSynchronize(c->mirrorStream));
    CK(cudaEventRecord(c->tm1, c->stream));
    CK(cudaEventSynchronize(c->tm1));
    float f = 0;
    CK(cudaEventElapsedTime(&f, c->tm0, c->tm1));
    *ms = f;
    return X265CU_OK;
}

/* synchronises: the job counts are those that passed their condition on the device */
int x265cu_get_counters(x265cu_ctx* c, x265cu_counters* o)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    int st = x265cu_sync(c);
    if (st) return st;
    unsigned long long ex[2] = { 0, 0 };
    CK(cudaMemcpy(ex, c->d_executed, sizeof(ex), cudaMemcpyDeviceToHost));
    c->counters.search_jobs = ex[0]; c->counters.cost_jobs = ex[1];
    *o = c->counters;
    return X265CU_OK;
}

int x265cu_profile_enable(x265cu_ctx* c, int32_t on)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    resolveProfile(c);
    c->profile = on != 0;
    return X265CU_OK;
}

int x265cu_profile_get_busy(x265cu_ctx* c, double busy[X265CU_K_COUNT])
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    resolveProfile(c);
    for (int i = 0; i < X265CU_K_COUNT; i++) busy[i] = c->profBusy[i];
    return X265CU_OK;
}

int x265cu_profile_get(x265cu_ctx* c, double ms[X265CU_K_COUNT], uint64_t launches[X265CU_K_COUNT], int32_t reset)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    resolveProfile(c);
    for (int i = 0; i < X265CU_K_COUNT; i++) { ms[i] = c->profMs[i]; launches[i] = c->profN[i]; }
    if (reset) { memset(c->profMs, 0, sizeof(c->profMs)); memset(c->profN, 0, sizeof(c->profN)); memset(c->profBusy, 0, sizeof(c->profBusy)); }
    return X265CU_OK;
}

int x265cu_frame_upload(x265cu_ctx* c, int32_t slot, const void* y, const void* u, const void* v, int32_t sy, int32_t sc)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!c || !slotOk(c, slot) || !y) return X265CU_ERR_BAD_ARG;
    if (c->cfg.hist_stats && (!u || !v))
    {
        /* --hist-scenecut reads the chroma planes (slicetype.cpp:1560-1640); on a 4:0:0 picture the reference dereferences NULL there */
        snprintf(c->err, sizeof(c->err), "hist-scenecut statistics need chroma planes (4:0:0 picture)");
        return X265CU_ERR_BAD_ARG;
    }
    return DISPATCH(uploadT, c, slot, y, u, v, sy, sc);
}

int x265cu_frame_ready(x265cu_ctx* c, int32_t slot)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    if (!c || !slotOk(c, slot)) return X265CU_ERR_BAD_ARG;
    /* test hook: every frame "still in flight", so that callers exercise the paths they take behind a busy GPU
     * (Lookahead::verifyWeights: weights assumed, settled later, searches redone) on sequences too short to get there */
    static const bool never = getenv("X265CU_FRAME_READY_NEVER") != NULL;
    if (never) return 0;
    const cudaError_t e = cudaEventQuery(c->slotConsumed[slot]);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    cudaOk(c, e, "cudaEventQuery");
    return X265CU_ERR_CUDA;
}

int x265cu_frame_hist_get(x265cu_ctx* c, int32_t slot, x265cu_hist_stats* out)
{
    if (!c || !out) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    if (!slotOk(c, slot) || !c->cfg.hist_stats) return X265CU_ERR_BAD_ARG;
    CK(cudaEventSynchronize(c->slotConsumed[slot]));
    memcpy(out, c->h_hist + slot, sizeof(*out));
    c->counters.d2h_bytes += sizeof(*out);
    return X265CU_OK;
}

int x265cu_frame_stats_get(x265cu_ctx* c, const int32_t* slots, int32_t n, x265cu_frame_stats* out)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    HostTimer ht(HT_STATS_WAIT);
    /* the statistics were published into mapped host memory by the frame's own pre-lookahead: wait for that only */
    for (int i = 0; i < n; i++)
    {
        if (!slotOk(c, slots[i])) return X265CU_ERR_BAD_ARG;
        CK(cudaEventSynchronize(c->slotConsumed[slots[i]]));
        const FrameStatsDev* s = c->h_slotStats + slots[i];
        out[i].cost_est = s->costEst; out[i].cost_est_aq = s->costEstAq;
        for (int k = 0; k < 3; k++) { out[i].wp_ssd[k] = s->wp_ssd[k]; out[i].wp_sum[k] = s->wp_sum[k]; }
        out[i].frame_variance = c->cfg.fade_stats ? s->frameVariance : 0;
    }
    c->counters.d2h_bytes += (n > 0 ? n : 0) * sizeof(FrameStatsDev);
    return X265CU_OK;
}

int x265cu_search_batch(x265cu_ctx* c, const x265cu_search_job* jobs, int32_t n)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!c || (n > 0 && !jobs)) return X265CU_ERR_BAD_ARG;
    if (n <= 0) return X265CU_OK;
    return DISPATCH(searchBatchT, c, jobs, n);
}

int x265cu_search_flags_get(x265cu_ctx* c, const int32_t* slots, const int32_t* stores, int32_t n, int32_t* flags)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (n <= 0) return X265CU_OK;
    std::vector<const void*> srcs(n);
    for (int i = 0; i < n; i++)
    {
        if (!slotOk(c, slots[i]) || stores[i] < 0 || stores[i] >= c->geom.n_mv_stores) return X265CU_ERR_BAD_ARG;
        /* (sharded stream: the flag of a frame another rank owns arrives with the exchange at the end of the batch) */
        int st = streamWaitBatch(c, c->gatherStream, c->mvWriter[(size_t)slots[i] * c->geom.n_mv_stores + stores[i]], c->nranks <= 1);
        if (st) return st;
        srcs[i] = mvStorePtr(c, slots[i], stores[i]) + (size_t)c->g.ncu * 8;
    }
    return gatherSmall(c, srcs, 1, flags);
}

int x265cu_cost_batch(x265cu_ctx* c, const x265cu_cost_job* jobs, int32_t n)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!c || (n > 0 && !jobs)) return X265CU_ERR_BAD_ARG;
    if (n <= 0) return X265CU_OK;
    return DISPATCH(costBatchT, c, jobs, n);
}

int x265cu_cost_results_get(x265cu_ctx* c, const int32_t* slots, const int32_t* outs, int32_t n, x265cu_cost_result* res)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (n <= 0) return X265CU_OK;
    std::vector<const void*> srcs(n);
    for (int i = 0; i < n; i++)
    {
        if (!slotOk(c, slots[i]) || outs[i] < 0 || outs[i] >= c->geom.n_cost_stores) return X265CU_ERR_BAD_ARG;
        int st = outs[i] < 2 ? X265CU_OK : streamWaitBatch(c, c->gatherStream, c->costWriter[(size_t)slots[i] * c->geom.n_cost_stores + outs[i]], false);
        if (st) return st;
        srcs[i] = costStorePtr(c, slots[i], outs[i]) + c->lay.costResOff;
    }
    std::vector<CostResultDev> tmp(n);
    int st = gatherSmall(c, srcs, (int)(sizeof(CostResultDev) / 4), &tmp[0]);
    if (st) return st;
    for (int i = 0; i < n; i++)
    {
        res[i].cost_est = tmp[i].costEst; res[i].cost_est_aq = tmp[i].costEstAq; res[i].intra_mbs = tmp[i].intraMbs; res[i].reserved = 0;
    }
    return X265CU_OK;
}

int x265cu_weight_cost_batch(x265cu_ctx* c, const x265cu_wcost_job* jobs, int32_t n, uint32_t* costs)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!c || (n > 0 && (!jobs || !costs))) return X265CU_ERR_BAD_ARG;
    if (n <= 0) return X265CU_OK;
    return DISPATCH(weightCostT, c, jobs, n, costs);
}

int x265cu_cutree_reset(x265cu_ctx* c, int32_t slot)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!slotOk(c, slot)) return X265CU_ERR_BAD_ARG;
    int st = mainWaitPre(c, slot);
    if (st) return st;
    CK(cudaMemsetAsync(c->slots[slot] + c->lay.propagate, 0, (size_t)c->g.ncu * 4, c->stream));
    return X265CU_OK;
}

int x265cu_cutree_propagate(x265cu_ctx* c, int32_t bs, int32_t p0s, int32_t p1s, int32_t cost_store, int32_t l0, int32_t l1,
                            int32_t referenced, int32_t bipred_weight, double fps_factor)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    if (!slotOk(c, bs) || !slotOk(c, p0s) || !slotOk(c, p1s) || cost_store < 2 || cost_store >= c->geom.n_cost_stores ||
        l0 < 0 || l0 >= c->geom.n_mv_stores || l1 >= c->geom.n_mv_stores)
        return X265CU_ERR_BAD_ARG;
    const SlotLayout& L = c->lay;
    HostTimer ht(HT_CUTREE);
    int st = mainWaitCost(c, bs, cost_store);
    if (!st) st = mainWaitMv(c, bs, l0);
    if (!st) st = mainWaitMv(c, bs, l1);
    if (!st) st = mainWaitPre(c, bs);
    if (!st) st = mainWaitPre(c, p0s);
    if (!st) st = mainWaitPre(c, p1s);
    if (st) return st;
    const int* mv0 = (const int*)mvStorePtr(c, bs, l0);
    const int* mv1 = l1 >= 0 ? (const int*)mvStorePtr(c, bs, l1) : mv0;
    if (!referenced)
    {
        /* gathered; launched together with its neighbours by the next call that is not one of these */
        if (c->ctRingPos + 1 > LA_CT_RING)
        {
            flushCutree(c);
            CK(cudaStreamSynchronize(c->stream));       /* the ring wraps: everything that read it has run */
            c->ctRingPos = 0;
        }
        CutreeJobDev& J = c->h_ctJobs[c->ctRingPos++];
        J.intraCost = slotPtr<int>(c, bs, L.intraCost); J.lowresCosts = (const unsigned short*)costStorePtr(c, bs, cost_store);
        J.invQ = slotInvQ(c, bs); J.mv0 = mv0; J.mv1 = mv1;
        J.ref0 = slotPtr<int>(c, p0s, L.propagate); J.ref1 = slotPtr<int>(c, p1s, L.propagate);
        J.self = slotPtr<int>(c, bs, L.propagate);
        J.bipredWeight = bipred_weight; J.pad = 0; J.fpsFactor = fps_factor;
        c->ctPending++;
        return X265CU_OK;
    }
    flushCutree(c);
    Prof pr(c, X265CU_K_CUTREE, 1);
    cutree_propagate_kernel<<<(c->g.ncu + 255) / 256, 256, 0, c->stream>>>(
        c->g, slotPtr<int>(c, bs, L.intraCost), (const unsigned short*)costStorePtr(c, bs, cost_store), slotInvQ(c, bs),
        mv0, mv1, slotPtr<int>(c, bs, L.propagate), slotPtr<int>(c, p0s, L.propagate),
        slotPtr<int>(c, p1s, L.propagate), bipred_weight, fps_factor);
    CK(cudaGetLastError());
    return X265CU_OK;
}

int x265cu_cutree_finish(x265cu_ctx* c, int32_t slot, int32_t fps_fix8, double weightdelta, double strength)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!slotOk(c, slot)) return X265CU_ERR_BAD_ARG;
    const SlotLayout& L = c->lay;
    int st = mainWaitPre(c, slot);
    if (st) return st;
    Prof pr(c, X265CU_K_CUTREE, 1);
    cutree_finish_kernel<<<(c->g.ncu + 255) / 256, 256, 0, c->stream>>>(
        c->g, slotPtr<int>(c, slot, L.intraCost), slotInvQ(c, slot), slotPtr<int>(c, slot, L.propagate),
        slotPtr<double>(c, slot, L.qpAq), slotPtr<double>(c, slot, L.qpCuTree), fps_fix8, weightdelta, strength);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->slotMainWrote[slot], c->stream));
    c->slotMainWroteSet[slot] = 1;
    return X265CU_OK;
}

int x265cu_cost_recalc(x265cu_ctx* c, int32_t slot, int32_t cost_store, int32_t use_cutree, int64_t* score, int32_t* rows)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!slotOk(c, slot) || cost_store < 0 || cost_store >= c->geom.n_cost_stores || cost_store == 1 || !score) return X265CU_ERR_BAD_ARG;
    const SlotLayout& L = c->lay;
    const Geom& g = c->g;
    const unsigned short* costs; int* rs;
    if (cost_store == 0) { costs = slotPtr<unsigned short>(c, slot, L.lowresCosts00); rs = slotPtr<int>(c, slot, L.rowSatds00); }
    else { char* cs = costStorePtr(c, slot, cost_store); costs = (const unsigned short*)cs; rs = (int*)(cs + L.costRowOff); }
    int st = mainWaitCost(c, slot, cost_store);
    if (!st) st = mainWaitPre(c, slot);
    if (st) return st;
    st = ensureDev(c, &c->d_results, &c->resultsCap, 8);
    if (st) return st;
    CK(cudaMemsetAsync(rs, 0, (size_t)g.bh * 4, c->stream));
    CK(cudaMemsetAsync(c->d_results, 0, 8, c->stream));
    {
        Prof pr(c, X265CU_K_CUTREE, 1);
        cost_recalc_kernel<<<(g.ncu + 255) / 256, 256, 0, c->stream>>>(g, costs, slotPtr<double>(c, slot, use_cutree ? L.qpCuTree : L.qpAq),
                                                                      rs, (unsigned long long*)c->d_results);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->slotMainWrote[slot], c->stream));
    c->slotMainWroteSet[slot] = 1;
    st = ensureMapped(c, 8 + (size_t)g.bh * 4);
    if (st) return st;
    publish_kernel<<<1, 32, 0, c->stream>>>((const unsigned*)c->d_results, (unsigned*)c->d_mapped, 2);
    if (rows) publish_kernel<<<1, 128, 0, c->stream>>>((const unsigned*)rs, (unsigned*)(c->d_mapped + 8), g.bh);
    c->counters.kernel_launches += 1 + (rows != NULL);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    *score = *(const long long*)c->h_mapped;
    if (rows) memcpy(rows, c->h_mapped + 8, (size_t)g.bh * 4);
    c->counters.d2h_bytes += 8 + (rows ? (size_t)g.bh * 4 : 0);
    return X265CU_OK;
}

static int d2h(x265cu_ctx* c, void* dst, const void* src, size_t bytes)
{
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    c->counters.d2h_bytes += bytes;
    return X265CU_OK;
}

int x265cu_vbv_row_costs(x265cu_ctx* c, int32_t slot, int32_t cost_store, int32_t qp_source, int32_t ctu_rows_lowres,
                         int32_t pir_start, int32_t pir_end, int32_t n_rows, uint32_t* satd, uint32_t* intra,
                         uint16_t* cost_for_rc, int32_t* intra_scaled)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    const Geom& g = c->g;
    const SlotLayout& L = c->lay;
    if (!slotOk(c, slot) || cost_store < 0 || cost_store >= c->geom.n_cost_stores || cost_store == 1 || ctu_rows_lowres < 1 ||
        qp_source < 0 || qp_source > 2 || n_rows < (g.bh + ctu_rows_lowres - 1) / ctu_rows_lowres)
        return X265CU_ERR_BAD_ARG;
    const unsigned short* costs = cost_store == 0 ? slotPtr<unsigned short>(c, slot, L.lowresCosts00)
                                                  : (const unsigned short*)costStorePtr(c, slot, cost_store);
    int st = mainWaitCost(c, slot, cost_store);
    if (!st) st = mainWaitPre(c, slot);
    /* scratch: row sums (2 x n_rows u32) | scaled costs (ncu u16) | scaled intra costs (ncu i32) */
    const size_t offCost = alignUp((size_t)n_rows * 8, 256), offIntra = offCost + alignUp((size_t)g.ncu * 2, 256);
    if (!st) st = ensureDev(c, &c->d_results, &c->resultsCap, offIntra + (size_t)g.ncu * 4);
    if (st) return st;
    CK(cudaMemsetAsync(c->d_results, 0, (size_t)n_rows * 8, c->stream));
    {
        Prof pr(c, X265CU_K_CUTREE, 1);
        const double* qp = qp_source == 0 || !c->cfg.need_aq ? NULL : slotPtr<double>(c, slot, qp_source == 2 ? L.qpCuTree : L.qpAq);
        vbv_rows_kernel<<<(g.ncu + 255) / 256, 256, 0, c->stream>>>(g, costs, slotPtr<int>(c, slot, L.intraCost), qp, ctu_rows_lowres,
                                                                   pir_start, pir_end, (unsigned short*)(c->d_results + offCost),
                                                                   (int*)(c->d_results + offIntra), (unsigned*)c->d_results,
                                                                   (unsigned*)c->d_results + n_rows);
    }
    CK(cudaGetLastError());
    if (satd) st = d2h(c, satd, c->d_results, (size_t)n_rows * 4);
    if (intra && !st) st = d2h(c, intra, c->d_results + (size_t)n_rows * 4, (size_t)n_rows * 4);
    if (cost_for_rc && !st) st = d2h(c, cost_for_rc, c->d_results + offCost, (size_t)g.ncu * 2);
    if (intra_scaled && !st) st = d2h(c, intra_scaled, c->d_results + offIntra, (size_t)g.ncu * 4);
    if (st) return st;
    CK(cudaStreamSynchronize(c->stream));
    return X265CU_OK;
}

__global__ void clamp_u16_kernel(const int* __restrict__ src, unsigned short* __restrict__ dst, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (unsigned short)min(max(src[i], 0), 65535);
}

__global__ void unpack_mv_kernel(const int* __restrict__ src, int* __restrict__ dst, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const int p = src[i]; dst[2 * i] = (int)(short)(p & 0xffff); dst[2 * i + 1] = p >> 16; }
}

struct UnpackTab { const int* src[X265CU_MIRROR_MAX_MV]; int* dst[X265CU_MIRROR_MAX_MV]; };
__global__ void __launch_bounds__(256) unpack_mvs_kernel(UnpackTab t, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = t.src[blockIdx.y][i];
    int2 v; v.x = (int)(short)(p & 0xffff); v.y = p >> 16;
    ((int2*)t.dst[blockIdx.y])[i] = v;
}

/* The small arrays of a mirror request in ONE launch: the kernel stores straight into the caller's page-locked arrays (registered
 * host memory is mapped into the device's address space), so a request costs one kernel launch instead of ~30 cudaMemcpyAsync
 * calls of 130-260 KB each -- measured 0.5-0.8 ms of host time per decided frame, the largest single item of the end-to-end
 * run.  kind 0: copy 32-bit words; kind 1: packed (int16, int16) MVs -> (int32, int32) pairs (Lowres::lowresMvs). */
struct MirrorSeg { const unsigned* src; unsigned* dst; unsigned words; unsigned kind; };
enum { LA_MIRROR_SEGS = X265CU_MIRROR_MAX_MV + 8 };
struct MirrorTab { MirrorSeg s[LA_MIRROR_SEGS]; };
__global__ void __launch_bounds__(256) mirror_scatter_kernel(MirrorTab t)
{
    const MirrorSeg sg = t.s[blockIdx.y];
    if (sg.kind == 0)
    {
        /* 16 bytes per thread where source, destination and length allow (all device arrays are 256-byte aligned) */
        const unsigned n4 = ((((size_t)sg.dst | (size_t)sg.src) & 15) == 0) ? sg.words >> 2 : 0;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
            ((uint4*)sg.dst)[i] = ((const uint4*)sg.src)[i];
        for (unsigned i = n4 * 4 + blockIdx.x * blockDim.x + threadIdx.x; i < sg.words; i += gridDim.x * blockDim.x)
            sg.dst[i] = sg.src[i];
    }
    else
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < sg.words; i += gridDim.x * blockDim.x)
        {
            const int p = (int)sg.src[i];
            uint2 v; v.x = (unsigned)(int)(short)(p & 0xffff); v.y = (unsigned)(p >> 16);
            ((uint2*)sg.dst)[i] = v;
        }
}

/* device address of a host pointer inside a page-locked (registered) range, or NULL; looked up once per pointer */
static void* mappedPtr(x265cu_ctx* c, void* host)
{
    std::map<void*, void*>::iterator it = c->mappedCache.find(host);
    if (it != c->mappedCache.end()) return it->second;
    void* dev = NULL;
    if (cudaHostGetDevicePointer(&dev, host, 0) != cudaSuccess) { cudaGetLastError(); dev = NULL; }
    c->mappedCache[host] = dev;
    return dev;
}

int x265cu_mirror_enqueue(x265cu_ctx* c, int32_t slot, const x265cu_mirror_request* q, int64_t* ticket)
{
    if (!c || !q) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    const Geom& g = c->g;
    const SlotLayout& L = c->lay;
    if (!slotOk(c, slot) || q->n_mv < 0 || q->n_mv > X265CU_MIRROR_MAX_MV || q->cost_store == 1 || q->cost_store >= c->geom.n_cost_stores)
        return X265CU_ERR_BAD_ARG;
    x265cu_ctx::MirrorEntry& e = c->mirror[c->nextMirror % X265CU_MIRROR_RING];
    if (e.ticket >= 0) { HostTimer ht(HT_MIRROR_RINGWAIT); CK(cudaEventSynchronize(e.done)); }  /* the ring wrapped: that request's scratch is free again */
    HostTimer htRest(HT_MIRROR_REST);
    const cudaStream_t ms = c->mirrorStream;
    /* behind the main-stream kernels that wrote what is copied here: the last cuTreeFinish (qpCuTreeOffset) / synchronous cost
     * recalculation (rowSatds) enqueued for THIS slot -- not behind the whole main stream: by the time the encoder takes a
     * decided frame the next decision's cuTree is already queued there, waiting for the newest batch, and a mirror ordered
     * behind it stalled the caller (and with it the feed of new pictures) for a whole batch */
    if (c->slotMainWroteSet[slot]) CK(cudaStreamWaitEvent(ms, c->slotMainWrote[slot], 0));
    /* ... the frame's own pre-lookahead and the batches that wrote the stores asked for */
    CK(cudaStreamWaitEvent(ms, c->slotConsumed[slot], 0));
    int st = X265CU_OK;
    long long waitedFor[X265CU_MIRROR_MAX_MV]; int nWaited = 0;     /* most stores of a frame were written by the same few batches */
    for (int i = 0; i < q->n_mv && !st; i++)
    {
        if (q->mv_store[i] < 0 || q->mv_store[i] >= c->geom.n_mv_stores || !q->mv_dst[i]) return X265CU_ERR_BAD_ARG;
        const long long w = c->mvWriter[(size_t)slot * c->geom.n_mv_stores + q->mv_store[i]];
        bool seen = false;
        for (int k = 0; k < nWaited; k++) seen |= waitedFor[k] == w;
        if (seen) continue;
        waitedFor[nWaited++] = w;
        st = streamWaitBatch(c, ms, w, c->nranks <= 1);
    }
    if (!st && q->cost_store >= 2) st = streamWaitBatch(c, ms, c->costWriter[(size_t)slot * c->geom.n_cost_stores + q->cost_store], false);
    if (st) return st;
    /* fast path: every small destination array is page-locked => one scatter kernel writes them all */
    bool direct = getenv("X265CU_MIRROR_MEMCPY") == NULL;
    MirrorTab tab; int nseg = 0; size_t segBytes = 0; unsigned maxWords = 0;
    if (direct)
    {
        struct Add
        {
            static bool seg(x265cu_ctx* c, MirrorTab& t, int& n, size_t& bytes, unsigned& maxWords, const void* src, void* hostDst, size_t words, unsigned kind)
            {
                if (!hostDst) return true;
                void* d = mappedPtr(c, hostDst);
                if (!d || n >= LA_MIRROR_SEGS) return false;
                t.s[n].src = (const unsigned*)src; t.s[n].dst = (unsigned*)d; t.s[n].words = (unsigned)words; t.s[n].kind = kind;
                n++; bytes += words * (kind ? 8 : 4); maxWords = std::max(maxWords, (unsigned)words);
                return true;
            }
        };
        direct = Add::seg(c, tab, nseg, segBytes, maxWords, c->slots[slot] + L.intraCost, q->intra_cost, g.ncu, 0) &&
                 Add::seg(c, tab, nseg, segBytes, maxWords, c->slots[slot] + L.qpAq, q->qp_aq_offset, (size_t)g.ncuFull * 2, 0) &&
                 Add::seg(c, tab, nseg, segBytes, maxWords, c->slots[slot] + L.qpCuTree, q->qp_cutree_offset, (size_t)g.ncuFull * 2, 0) &&
                 (!c->cfg.need_aq || Add::seg(c, tab, nseg, segBytes, maxWords, c->slots[slot] + L.invQ, q->inv_qscale_factor, g.ncuFull, 0));
        for (int i = 0; direct && i < q->n_mv; i++)
            direct = Add::seg(c, tab, nseg, segBytes, maxWords, mvStorePtr(c, slot, q->mv_store[i]), q->mv_dst[i], g.ncu, 1);
        if (direct && q->cost_store >= 0 && (q->lowres_costs || q->row_satds))
        {
            const char* cs = q->cost_store == 0 ? NULL : costStorePtr(c, slot, q->cost_store);
            /* lowresCosts are 16-bit: whole words only when the block count is even, else that one array goes the slow way */
            if (q->lowres_costs && (g.ncu & 1)) direct = false;
            direct = direct && Add::seg(c, tab, nseg, segBytes, maxWords, cs ? cs : c->slots[slot] + L.lowresCosts00, q->lowres_costs, g.ncu / 2, 0) &&
                     Add::seg(c, tab, nseg, segBytes, maxWords, cs ? cs + L.costRowOff : c->slots[slot] + L.rowSatds00, q->row_satds, g.bh, 0);
        }
    }
    const size_t planeBytes = q->planes ? alignUp((size_t)(4 * g.planeSize) * c->bpp, 256) : 0;
    const size_t mvBytes = alignUp((size_t)g.ncu * 8, 256);
    const size_t need = planeBytes + (direct ? 0 : mvBytes * q->n_mv) + 256;
    if (e.cap < need)
    {
        /* cudaFree / cudaMalloc wait for the whole device: with several batches queued that is tens of milliseconds per call
         * (measured: 23 calls, 175 - 1100 ms per 300-frame step while the ring entries grew one request at a time).  So the
         * first request that needs scratch sizes EVERY ring entry for the largest request there can be, once. */
        HostTimer ht(HT_MIRROR_MALLOC);
        const size_t full = alignUp((size_t)(4 * g.planeSize) * c->bpp, 256) + mvBytes * X265CU_MIRROR_MAX_MV + 256;
        for (int i = 0; i < X265CU_MIRROR_RING; i++)
        {
            x265cu_ctx::MirrorEntry& m = c->mirror[i];
            if (m.cap >= full) continue;
            if (m.ticket >= 0) CK(cudaEventSynchronize(m.done));
            if (m.scratch) cudaFree(m.scratch);
            m.scratch = NULL; m.cap = 0;
            CK(cudaMalloc((void**)&m.scratch, full));
            m.cap = full;
        }
    }
#define MIRROR_D2H(dst, src, bytes) do { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ms)); c->counters.d2h_bytes += (bytes); } while (0)
    if (direct)
    {
        if (q->inv_qscale_factor && !c->cfg.need_aq) for (int i = 0; i < g.ncuFull; i++) q->inv_qscale_factor[i] = 256;
        if (nseg)
        {
            for (int i = nseg; i < LA_MIRROR_SEGS; i++) tab.s[i].words = 0;
            const unsigned gx = std::max(1u, std::min(64u, (maxWords / 4 + 255) / 256));
            mirror_scatter_kernel<<<dim3(gx, nseg), 256, 0, ms>>>(tab);
            c->counters.kernel_launches++;
            c->counters.d2h_bytes += segBytes;
            CK(cudaGetLastError());
        }
    }
    else
    {
    if (q->intra_cost) MIRROR_D2H(q->intra_cost, c->slots[slot] + L.intraCost, (size_t)g.ncu * 4);
    if (q->qp_aq_offset) MIRROR_D2H(q->qp_aq_offset, c->slots[slot] + L.qpAq, (size_t)g.ncuFull * 8);
    if (q->qp_cutree_offset) MIRROR_D2H(q->qp_cutree_offset, c->slots[slot] + L.qpCuTree, (size_t)g.ncuFull * 8);
    if (q->inv_qscale_factor)
    {
        if (c->cfg.need_aq) MIRROR_D2H(q->inv_qscale_factor, c->slots[slot] + L.invQ, (size_t)g.ncuFull * 4);
        else for (int i = 0; i < g.ncuFull; i++) q->inv_qscale_factor[i] = 256;
    }
    if (q->n_mv)
    {
        UnpackTab tab;
        for (int i = 0; i < q->n_mv; i++)
        {
            tab.src[i] = (const int*)mvStorePtr(c, slot, q->mv_store[i]);
            tab.dst[i] = (int*)(e.scratch + planeBytes + mvBytes * i);
        }
        unpack_mvs_kernel<<<dim3((g.ncu + 255) / 256, q->n_mv), 256, 0, ms>>>(tab, g.ncu);
        c->counters.kernel_launches++;
        CK(cudaGetLastError());
        for (int i = 0; i < q->n_mv; i++) MIRROR_D2H(q->mv_dst[i], tab.dst[i], (size_t)g.ncu * 8);
    }
    if (q->cost_store >= 0 && (q->lowres_costs || q->row_satds))
    {
        const char* cs = q->cost_store == 0 ? NULL : costStorePtr(c, slot, q->cost_store);
        if (q->lowres_costs) MIRROR_D2H(q->lowres_costs, cs ? cs : c->slots[slot] + L.lowresCosts00, (size_t)g.ncu * 2);
        if (q->row_satds) MIRROR_D2H(q->row_satds, cs ? cs + L.costRowOff : c->slots[slot] + L.rowSatds00, (size_t)g.bh * 4);
    }
    }
    if (q->planes)
    {
        const unsigned blocks = (unsigned)((4 * g.planeSize + 255) / 256);
        if (c->bpp == 1) detile_kernel<uint8_t><<<blocks, 256, 0, ms>>>(g, slotPtr<uint8_t>(c, slot, L.planes), (uint8_t*)e.scratch);
        else detile_kernel<uint16_t><<<blocks, 256, 0, ms>>>(g, slotPtr<uint16_t>(c, slot, L.planes), (uint16_t*)e.scratch);
        c->counters.kernel_launches++;
        CK(cudaGetLastError());
        MIRROR_D2H(q->planes, e.scratch, (size_t)(4 * g.planeSize) * c->bpp);
    }
#undef MIRROR_D2H
    CK(cudaEventRecord(e.done, ms));
    CK(cudaEventRecord(c->slotMirrored[slot], ms));
    c->slotMirrorTouched[slot] = 1;
    e.ticket = c->nextMirror++;
    if (ticket) *ticket = e.ticket;
    return X265CU_OK;
}

int x265cu_mirror_wait(x265cu_ctx* c, int64_t ticket)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    if (ticket < 0 || ticket >= c->nextMirror) return X265CU_ERR_BAD_ARG;
    x265cu_ctx::MirrorEntry& e = c->mirror[ticket % X265CU_MIRROR_RING];
    if (e.ticket != ticket) return X265CU_OK;       /* the ring already moved past it: it was waited for then */
    CK(cudaEventSynchronize(e.done));
    return X265CU_OK;
}

int x265cu_cost_recalc_enqueue(x265cu_ctx* c, int32_t slot, int32_t cost_store, int32_t use_cutree)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!slotOk(c, slot) || cost_store < 0 || cost_store >= c->geom.n_cost_stores || cost_store == 1) return X265CU_ERR_BAD_ARG;
    const SlotLayout& L = c->lay;
    const Geom& g = c->g;
    const unsigned short* costs = cost_store == 0 ? slotPtr<unsigned short>(c, slot, L.lowresCosts00)
                                                  : (const unsigned short*)costStorePtr(c, slot, cost_store);
    int st = mainWaitCost(c, slot, cost_store);
    if (!st) st = mainWaitPre(c, slot);
    if (st) return st;
    char* scratch = c->d_recalc + (size_t)slot * c->recalcStride;
    CK(cudaMemsetAsync(scratch, 0, 8 + (size_t)g.bh * 4, c->stream));
    {
        Prof pr(c, X265CU_K_CUTREE, 1);
        cost_recalc_kernel<<<(g.ncu + 255) / 256, 256, 0, c->stream>>>(g, costs, slotPtr<double>(c, slot, use_cutree ? L.qpCuTree : L.qpAq),
                                                                      (int*)(scratch + 8), (unsigned long long*)scratch);
    }
    CK(cudaGetLastError());
    {
        void* hAlias = hostAlias(c->h_recalc + (size_t)slot * c->recalcStride);
        if (!hAlias) { snprintf(c->err, sizeof(c->err), "recalc result memory is not device-visible"); return X265CU_ERR_CUDA; }
        st = stageLaunch(c, c->stream, hAlias, scratch, alignUp(8 + (size_t)g.bh * 4, 16), NULL, NULL, 0, NULL, 0);
        if (st) return st;
    }
    CK(cudaEventRecord(c->slotRecalc[slot], c->stream));
    c->slotRecalcStore[slot] = cost_store;
    return X265CU_OK;
}

int x265cu_cost_recalc_get(x265cu_ctx* c, int32_t slot, int32_t cost_store, int64_t* score, int32_t* rows)
{
    if (!c || !score) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    if (!slotOk(c, slot) || c->slotRecalcStore[slot] != cost_store) return X265CU_ERR_BAD_ARG;
    HostTimer ht(HT_RECALC_GET);
    const SlotLayout& L = c->lay;
    const Geom& g = c->g;
    CK(cudaEventSynchronize(c->slotRecalc[slot]));
    const char* h = c->h_recalc + (size_t)slot * c->recalcStride;
    *score = *(const long long*)h;
    if (rows) memcpy(rows, h + 8, (size_t)g.bh * 4);
    c->counters.d2h_bytes += 8 + (size_t)g.bh * 4;
    /* the reference's call leaves the recalculated row sums in rowSatds[b - p0][p1 - b]: move them there now */
    int* rs = cost_store == 0 ? slotPtr<int>(c, slot, L.rowSatds00) : (int*)(costStorePtr(c, slot, cost_store) + L.costRowOff);
    CK(cudaMemcpyAsync(rs, c->d_recalc + (size_t)slot * c->recalcStride + 8, (size_t)g.bh * 4, cudaMemcpyDeviceToDevice, c->stream));
    c->slotMainTouched[slot] = 1;
    c->slotRecalcStore[slot] = -1;
    return X265CU_OK;
}

int x265cu_fetch_frame(x265cu_ctx* c, int32_t slot, const x265cu_frame_out* o)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!slotOk(c, slot) || !o) return X265CU_ERR_BAD_ARG;
    const SlotLayout& L = c->lay;
    const Geom& g = c->g;
    int st = mainWaitPre(c, slot);
    if (o->intra_cost && !st) st = d2h(c, o->intra_cost, c->slots[slot] + L.intraCost, (size_t)g.ncu * 4);
    if (o->intra_mode && !st) st = d2h(c, o->intra_mode, c->slots[slot] + L.intraMode, (size_t)g.ncu);
    if (o->qp_aq_offset && !st) st = d2h(c, o->qp_aq_offset, c->slots[slot] + L.qpAq, (size_t)g.ncuFull * 8);
    if (o->qp_cutree_offset && !st) st = d2h(c, o->qp_cutree_offset, c->slots[slot] + L.qpCuTree, (size_t)g.ncuFull * 8);
    if (o->inv_qscale_factor && !st)
    {
        if (c->cfg.need_aq) st = d2h(c, o->inv_qscale_factor, c->slots[slot] + L.invQ, (size_t)g.ncuFull * 4);
        else for (int i = 0; i < g.ncuFull; i++) o->inv_qscale_factor[i] = 256;   /* no AQ arrays: neutral scale */
    }
    if (o->propagate_cost && !st)
    {
        st = ensureDev(c, &c->d_results, &c->resultsCap, (size_t)g.ncu * 2);
        if (!st)
        {
            clamp_u16_kernel<<<(g.ncu + 255) / 256, 256, 0, c->stream>>>(slotPtr<int>(c, slot, L.propagate), (unsigned short*)c->d_results, g.ncu);
            c->counters.kernel_launches++;
            st = d2h(c, o->propagate_cost, c->d_results, (size_t)g.ncu * 2);
            if (!st && cudaStreamSynchronize(c->stream) != cudaSuccess) st = X265CU_ERR_CUDA;   /* d_results is reused below */
        }
    }
    if (o->planes && !st)
    {
        /* planes live tiled in HBM; the mirror is the reference's pitched Lowres::buffer[0..3] */
        const size_t bytes = (size_t)(4 * g.planeSize) * c->bpp;
        st = ensureDev(c, &c->d_results, &c->resultsCap, bytes);
        if (!st)
        {
            const unsigned blocks = (unsigned)((4 * g.planeSize + 255) / 256);
            if (c->bpp == 1) detile_kernel<uint8_t><<<blocks, 256, 0, c->stream>>>(g, slotPtr<uint8_t>(c, slot, L.planes), (uint8_t*)c->d_results);
            else detile_kernel<uint16_t><<<blocks, 256, 0, c->stream>>>(g, slotPtr<uint16_t>(c, slot, L.planes), (uint16_t*)c->d_results);
            c->counters.kernel_launches++;
            st = d2h(c, o->planes, c->d_results, bytes);
            if (!st && cudaStreamSynchronize(c->stream) != cudaSuccess) st = X265CU_ERR_CUDA;   /* d_results is reused below */
        }
    }
    if (o->lowres_costs00 && !st) st = d2h(c, o->lowres_costs00, c->slots[slot] + L.lowresCosts00, (size_t)g.ncu * 2);
    if (o->row_satds00 && !st) st = d2h(c, o->row_satds00, c->slots[slot] + L.rowSatds00, (size_t)g.bh * 4);
    if (st) return st;
    CK(cudaStreamSynchronize(c->stream));
    return X265CU_OK;
}

int x265cu_fetch_mvs(x265cu_ctx* c, int32_t slot, int32_t store, int32_t* mv, int32_t* cost)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!slotOk(c, slot) || store < 0 || store >= c->geom.n_mv_stores) return X265CU_ERR_BAD_ARG;
    const Geom& g = c->g;
    const int* st0 = (const int*)mvStorePtr(c, slot, store);
    int st = mainWaitMv(c, slot, store);
    if (st) return st;
    if (mv)
    {
        st = ensureDev(c, &c->d_results, &c->resultsCap, (size_t)g.ncu * 8);
        if (st) return st;
        unpack_mv_kernel<<<(g.ncu + 255) / 256, 256, 0, c->stream>>>(st0, (int*)c->d_results, g.ncu);
        c->counters.kernel_launches++;
        st = d2h(c, mv, c->d_results, (size_t)g.ncu * 8);
    }
    if (cost && !st) st = d2h(c, cost, st0 + g.ncu, (size_t)g.ncu * 4);
    if (st) return st;
    CK(cudaStreamSynchronize(c->stream));
    return X265CU_OK;
}

int x265cu_fetch_hme_mvs(x265cu_ctx* c, int32_t slot, int32_t store, int32_t* mv, int32_t* cost)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!c->cfg.hme || !slotOk(c, slot) || store < 0 || store >= c->geom.n_mv_stores) return X265CU_ERR_BAD_ARG;
    const int n4 = c->g4.ncu;
    const int* st0 = (const int*)(c->slots[slot] + c->lay.mvStores4 + (size_t)store * c->lay.mvStore4Stride);
    int st = mainWaitMv(c, slot, store);
    if (st) return st;
    if (mv)
    {
        st = ensureDev(c, &c->d_results, &c->resultsCap, (size_t)n4 * 8);
        if (st) return st;
        unpack_mv_kernel<<<(n4 + 255) / 256, 256, 0, c->stream>>>(st0, (int*)c->d_results, n4);
        c->counters.kernel_launches++;
        st = d2h(c, mv, c->d_results, (size_t)n4 * 8);
    }
    if (cost && !st) st = d2h(c, cost, st0 + n4, (size_t)n4 * 4);
    if (st) return st;
    CK(cudaStreamSynchronize(c->stream));
    return X265CU_OK;
}

int x265cu_fetch_costs(x265cu_ctx* c, int32_t slot, int32_t store, uint16_t* costs, int32_t* rows)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!slotOk(c, slot) || store < 2 || store >= c->geom.n_cost_stores) return X265CU_ERR_BAD_ARG;
    char* cs = costStorePtr(c, slot, store);
    int st = mainWaitCost(c, slot, store);
    if (st) return st;
    if (costs) st = d2h(c, costs, cs, (size_t)c->g.ncu * 2);
    if (rows && !st) st = d2h(c, rows, cs + c->lay.costRowOff, (size_t)c->g.bh * 4);
    if (st) return st;
    CK(cudaStreamSynchronize(c->stream));
    return X265CU_OK;
}

/* unit-test hook (not part of the drop-in surface): SAD / SATD of n packed 8x8 block pairs */
int x265cu_debug_block_metrics(x265cu_ctx* c, const void* a, const void* b, int32_t n, int32_t* sad, int32_t* satd)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!c || !a || !b || n <= 0) return X265CU_ERR_BAD_ARG;
    return DISPATCH(blockMetricsT, c, a, b, n, sad, satd);
}

/* unit-test hook: lowresQPelCost (SAD and SATD) of n (block index, quarter-pel MV) pairs, fenc vs ref slot */
int x265cu_debug_mc_metrics(x265cu_ctx* c, int32_t fenc_slot, int32_t ref_slot, const int32_t* cu_idx, const int32_t* mvs, int32_t n,
                            int32_t* sad, int32_t* satd)
{
    if (!c) return X265CU_ERR_BAD_ARG;
    DeviceScope deviceScope(c);
    flushCutree(c);
    if (!c || !slotOk(c, fenc_slot) || !slotOk(c, ref_slot) || n <= 0) return X265CU_ERR_BAD_ARG;
    int st = mainWaitPre(c, fenc_slot);
    if (!st) st = mainWaitPre(c, ref_slot);
    if (st) return st;
    return DISPATCH(mcMetricsT, c, fenc_slot, ref_slot, cu_idx, mvs, n, sad, satd);
}

} // extern "C"
