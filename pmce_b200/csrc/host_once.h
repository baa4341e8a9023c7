// Host-side helpers shared by the launchers: per-device one-time kernel configuration, per-device SM count, and the
// environment knobs. Nothing here is cached per PROCESS that CUDA keeps per DEVICE (cudaFuncSetAttribute is per device:
// an engine on a second GPU of the same process must configure its kernels again).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>

constexpr int PMCE_MAX_DEVICES = 64;

// cudaFuncAttributeMaxDynamicSharedMemorySize for `Kernel` on the CURRENT device, once per (kernel, device). Thread-safe:
// a racing second caller at worst sets the attribute twice.
template <auto Kernel>
static inline bool pmce_configure_smem(int bytes) {
    static std::atomic<unsigned long long> done{0};            // bit d = configured on device d
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (dev < 0 || dev >= PMCE_MAX_DEVICES) return cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
    const unsigned long long bit = 1ull << dev;
    if (done.load(std::memory_order_acquire) & bit) return true;
    if (cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
    done.fetch_or(bit, std::memory_order_release);
    return true;
}

// SM count of the current device (cached per device)
static inline int tc_num_sms() {
    static std::atomic<int> n[PMCE_MAX_DEVICES];
    int dev = 0;
    cudaGetDevice(&dev);
    const bool cacheable = dev >= 0 && dev < PMCE_MAX_DEVICES;
    int v = cacheable ? n[dev].load(std::memory_order_relaxed) : 0;
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        if (cacheable) n[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// Tuning knob that does NOT change results (tile width, pair on/off, fused on/off): read once per process.
static inline int pmce_env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}

// Which launches carry the programmatic-serialisation attribute: PMCE_PDL is a mask over the launch SCOPE the orchestration sets
// (bit 0: the pose lifter while the image-feature stream runs beside it, bit 1: the image-feature stream, bit 2: everything else -
// the decoder and stand-alone C-ABI calls). Results are bit-identical for every mask (tools/ab_graphs.py checks it).
// DEFAULT 0 (off) - measured on B200, B=64 (profiles/r2y_pdl_ab.txt): each stage ALONE gets faster with PDL (lifter 1606 -> 1577 us,
// GRU -> y[T//2] 555 -> 514, decoder 1230 -> 1165), but the whole two-stream forward does not (2704 vs 2704 us for mask 5, +1 % for
// mask 4, +3 % for mask 7; at B=256 +1..4 %): a PDL chain never leaves a gap, so the chained GRU steps on the high-priority side
// stream keep their 128 SMs until the chain ends and the lifter starves (mask 7 = the two streams run one after the other),
// while on the lifter's stream the gaps PDL closes were already being filled by the other stream's CTAs.
constexpr int PMCE_PDL_LIFTER = 0, PMCE_PDL_SIDE = 1, PMCE_PDL_REST = 2;
static inline int& pmce_pdl_scope() {
    static thread_local int scope = PMCE_PDL_REST;
    return scope;
}
struct PdlScope {          // RAII: launches issued by this thread inside the scope belong to `s`
    int prev;
    explicit PdlScope(int s) : prev(pmce_pdl_scope()) { pmce_pdl_scope() = s; }
    ~PdlScope() { pmce_pdl_scope() = prev; }
};
static inline bool pmce_pdl_enabled() {
    // read at every launch (not cached): tools/ab_graphs.py captures one CUDA graph per mask in ONE process and replays them
    // interleaved, the only A/B that survives the box-to-box and thermal drift of +-1.5 %
    const int mask = pmce_env_int("PMCE_PDL", 0) & 7;
    return (mask >> pmce_pdl_scope()) & 1;
}

// Launch with programmatic stream serialisation (common.cuh pdl_wait / pdl_trigger): the kernel MUST call pdl_wait() before its
// first access to memory other kernels produce or reuse. cluster > 0 adds a cluster dimension along x. Captured into CUDA graphs
// the dependency becomes a programmatic edge.
template <typename... KArgs, typename... Args>
static inline cudaError_t pmce_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cluster > 0) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pmce_pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Profiling knob that makes the kernels produce WRONG results on purpose (skip loads / math / stores to time the rest).
// Honoured only in a -DPMCE_PROFILING build (PMCE_B200_PROFILING=1 python -m pmce_b200.build -> libpmce_b200_prof.so); the
// production library ignores it, loudly, so a leaked variable cannot silently corrupt a forward.
static inline int pmce_profiling_knob(const char* name) {
    const char* s = getenv(name);
    if (!s || !atoi(s)) return 0;
#ifdef PMCE_PROFILING
    fprintf(stderr, "[pmce_b200] PROFILING BUILD: %s=%s is active - results are wrong by design\n", name, s);
    return atoi(s);
#else
    fprintf(stderr, "[pmce_b200] %s=%s ignored: profiling knobs need the -DPMCE_PROFILING build (PMCE_B200_PROFILING=1)\n", name, s);
    return 0;
#endif
}
