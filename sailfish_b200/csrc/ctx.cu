// ctx.cu -- context, streams, pinned host memory, the NCCL communicator (loaded at run time) and the two
// known-answer hooks (device XXH64, device digamma).
#include <dlfcn.h>

#include <cstring>

#include <cstdlib>

#include <cctype>
#include <cstdio>
#include <string>

#include <sched.h>

#include "common.cuh"

extern "C" int sfb200_version(void) { return 100; }

extern "C" int sfb200_ctx_create(int device, sfb200_ctx** out) {
    if (!out) return SFB200_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return SFB200_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return SFB200_ENODEV;
    sfb200_ctx* c = new sfb200_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return SFB200_ENODEV; }
    c->num_sms = prop.multiProcessorCount;
    c->coop = prop.cooperativeLaunch;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return SFB200_ECUDA; }
    c->stream = c->own_stream;
    // the mapping kernels read random 32-byte sectors (table slots, suffix entries, bitmap words): SFB200_L2_FETCH=32|64|128 sets the
    // granularity L2 fetches from HBM with (a device-wide hint; default: the driver's)
    if (const char* e = getenv("SFB200_L2_FETCH")) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)); cudaGetLastError(); }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    *out = c;
    return SFB200_OK;
}

// "0-13,56-69" -> CPU set; false when the text holds anything else
static bool parse_cpulist(const std::string& txt, cpu_set_t* set) {
    CPU_ZERO(set);
    size_t i = 0;
    int n = 0;
    while (i < txt.size()) {
        if (isspace((unsigned char)txt[i]) || txt[i] == ',') { ++i; continue; }
        if (!isdigit((unsigned char)txt[i])) return false;
        long a = 0, b;
        while (i < txt.size() && isdigit((unsigned char)txt[i])) a = a * 10 + (txt[i++] - '0');
        b = a;
        if (i < txt.size() && txt[i] == '-') {
            ++i; b = 0;
            if (i >= txt.size() || !isdigit((unsigned char)txt[i])) return false;
            while (i < txt.size() && isdigit((unsigned char)txt[i])) b = b * 10 + (txt[i++] - '0');
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++n; }
    }
    return n > 0;
}

static bool slurp(const std::string& path, std::string* out) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[4096];
    const size_t n = fread(buf, 1, sizeof(buf) - 1, f);
    fclose(f);
    out->assign(buf, n);
    return n > 0;
}

extern "C" int sfb200_bind_host_near_device(int device) {
    if (const char* e = getenv("SFB200_NO_BIND")) if (atoi(e) != 0) return 0;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return 0; }   // no such device: sfb200_ctx_create says so
    for (char* p = bus; *p; ++p) *p = (char)tolower((unsigned char)*p);                 // sysfs spells the address in lower case
    std::string txt;
    if (!slurp(std::string("/sys/bus/pci/devices/") + bus + "/numa_node", &txt)) return 0;
    const int node = atoi(txt.c_str());
    if (node < 0) return 0;                                                             // the platform does not say
    if (!slurp("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", &txt)) return 0;
    cpu_set_t want, have, both;
    if (!parse_cpulist(txt, &want)) return 0;
    if (sched_getaffinity(0, sizeof(have), &have) != 0) return SFB200_EINVAL;
    CPU_AND(&both, &want, &have);                                                       // never widen what the launcher allowed
    const int n = CPU_COUNT(&both);
    if (n == 0 || n == CPU_COUNT(&have)) return 0;
    if (sched_setaffinity(0, sizeof(both), &both) != 0) return SFB200_EINVAL;
    return n;
}

extern "C" void sfb200_ctx_destroy(sfb200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    sfb_map_state_free(c);
    sfb_em_extra_free(c);
    sfb_fastq_free(c);
    c->index.words.release(); c->index.txp_start.release(); c->index.txp_len.release();
    c->index.txp_end.release(); c->index.sa.release(); c->index.table.release(); c->index.mfilter.release();
    c->cls.start.release(); c->cls.len.release(); c->cls.lab.release(); c->cls.w.release(); c->cls.cnt.release();
    c->cls.perm.release(); c->cls.sgl_cls.release(); c->cls.sgl_tid.release(); c->cls.cnt_all.release();
    c->cls.single.release(); c->cls.active.release(); c->cls.part.release();
    c->em_alpha.release(); c->em_theta.release(); c->em_base.release(); c->em_ctl.release(); c->eff.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" const char* sfb200_last_error(const sfb200_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" int sfb200_ctx_set_stream(sfb200_ctx* c, void* s) {
    if (!c) return SFB200_EINVAL;
    c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
    return SFB200_OK;
}

extern "C" int sfb200_ctx_sync(sfb200_ctx* c) {
    if (!c) return SFB200_EINVAL;
    SFB_CUDA(c, cudaStreamSynchronize(c->stream));
    return SFB200_OK;
}

extern "C" uint64_t sfb200_launch_count(const sfb200_ctx* c) { return c ? c->launches : 0; }

extern "C" void* sfb200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void sfb200_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- NCCL, resolved with dlopen so that libsfb200 has no link-time dependency on it ------------------------------
namespace {
struct Id128 { char b[128]; };
struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /*ncclUniqueId by value: 128 bytes*/ Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;

bool nccl_load() {
    if (g_nccl.h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return false;
    g_nccl.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(g_nccl.h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<int (*)(void**, int, Id128, int)>(dlsym(g_nccl.h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(g_nccl.h, "ncclAllReduce"));
    g_nccl.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(g_nccl.h, "ncclAllGather"));
    g_nccl.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(g_nccl.h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(g_nccl.h, "ncclGetErrorString"));
    return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce;
}
constexpr int NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_SUM = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)
}  // namespace

extern "C" int sfb200_comm_unique_id(uint8_t id_out[128]) {
    if (!nccl_load()) return SFB200_ENCCL;
    Id128 id;
    std::memset(&id, 0, sizeof(id));
    if (g_nccl.GetUniqueId(&id) != 0) return SFB200_ENCCL;
    std::memcpy(id_out, &id, 128);
    return SFB200_OK;
}

extern "C" int sfb200_comm_init(sfb200_ctx* c, int n_ranks, int rank, const uint8_t id[128]) {
    if (!c || n_ranks < 1 || rank < 0 || rank >= n_ranks) return SFB200_EINVAL;
    if (n_ranks == 1) { c->n_ranks = 1; c->rank = 0; return SFB200_OK; }
    if (!nccl_load()) SFB_FAIL(c, SFB200_ENCCL, "libnccl.so.2 could not be loaded");
    Id128 uid;
    std::memcpy(&uid, id, 128);
    cudaSetDevice(c->device);
    void* comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, n_ranks, uid, rank);
    if (rc != 0) SFB_FAIL(c, SFB200_ENCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    c->comm = comm; c->n_ranks = n_ranks; c->rank = rank;
    return SFB200_OK;
}

int sfb_comm_allreduce_f64(sfb200_ctx* c, double* d, size_t n) {
    if (c->n_ranks <= 1) return SFB200_OK;
    const int rc = g_nccl.AllReduce(d, d, n, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream);
    if (rc != 0) SFB_FAIL(c, SFB200_ENCCL, "ncclAllReduce(f64) failed");
    return SFB200_OK;
}
// every rank contributes `bytes` bytes; recv holds n_ranks * bytes
int sfb_comm_allgather(sfb200_ctx* c, const void* send, void* recv, size_t bytes) {
    if (c->n_ranks <= 1) return SFB200_OK;
    if (!g_nccl.AllGather) SFB_FAIL(c, SFB200_ENCCL, "ncclAllGather not found");
    const int rc = g_nccl.AllGather(send, recv, bytes, /*ncclUint8*/ 1, c->comm, c->stream);
    if (rc != 0) SFB_FAIL(c, SFB200_ENCCL, "ncclAllGather failed");
    return SFB200_OK;
}
int sfb_comm_allreduce_u64(sfb200_ctx* c, unsigned long long* d, size_t n) {
    if (c->n_ranks <= 1) return SFB200_OK;
    const int rc = g_nccl.AllReduce(d, d, n, NCCL_UINT64, NCCL_SUM, c->comm, c->stream);
    if (rc != 0) SFB_FAIL(c, SFB200_ENCCL, "ncclAllReduce(u64) failed");
    return SFB200_OK;
}

// ---- known-answer hooks ---------------------------------------------------------------------------------------------
__global__ void k_xxh64_msgs(const uint8_t* __restrict__ data, const uint64_t* __restrict__ off, uint64_t n, uint64_t seed,
                             uint64_t* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(data + off[i]);
    const uint32_t nw = static_cast<uint32_t>((off[i + 1] - off[i]) / 4);
    out[i] = xxh64_words([&](uint32_t j) { return w[j]; }, nw, seed);
}

extern "C" int sfb200_xxh64_device(sfb200_ctx* c, const uint8_t* data, const uint64_t* off, uint64_t n, uint64_t seed, uint64_t* out) {
    if (!c || !off || !out) return SFB200_EINVAL;
    if (n == 0) return SFB200_OK;
    const uint64_t total = off[n];
    for (uint64_t i = 0; i <= n; ++i) if (off[i] % 4) SFB_FAIL(c, SFB200_EINVAL, "message offsets must be multiples of 4");
    cudaSetDevice(c->device);
    DevBuf<uint8_t> d_data; DevBuf<uint64_t> d_off, d_out;
    DevBufScope<DevBuf<uint8_t>, DevBuf<uint64_t>, DevBuf<uint64_t>> scope(d_data, d_off, d_out);
    SFB_CUDA(c, d_data.reserve(total + 4));
    SFB_CUDA(c, d_off.reserve(n + 1));
    SFB_CUDA(c, d_out.reserve(n));
    if (total) SFB_CUDA(c, cudaMemcpyAsync(d_data.p, data, total, cudaMemcpyHostToDevice, c->stream));
    SFB_CUDA(c, cudaMemcpyAsync(d_off.p, off, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    k_xxh64_msgs<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(d_data.p, d_off.p, n, seed, d_out.p);
    c->launches++;
    SFB_CUDA(c, cudaGetLastError());
    SFB_CUDA(c, cudaMemcpyAsync(out, d_out.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    SFB_CUDA(c, cudaStreamSynchronize(c->stream));
    return SFB200_OK;
}

__global__ void k_digamma(const double* __restrict__ x, uint64_t n, double* __restrict__ out, int exp_form) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = exp_form ? sfb_exp_digamma(x[i]) : sfb_digamma(x[i]);
}

static int digamma_device(sfb200_ctx* c, const double* x, uint64_t n, double* out, int exp_form);
extern "C" int sfb200_digamma_device(sfb200_ctx* c, const double* x, uint64_t n, double* out) { return digamma_device(c, x, n, out, 0); }
extern "C" int sfb200_exp_digamma_device(sfb200_ctx* c, const double* x, uint64_t n, double* out) { return digamma_device(c, x, n, out, 1); }

static int digamma_device(sfb200_ctx* c, const double* x, uint64_t n, double* out, int exp_form) {
    if (!c || !x || !out) return SFB200_EINVAL;
    if (n == 0) return SFB200_OK;
    cudaSetDevice(c->device);
    DevBuf<double> d_x, d_o;
    DevBufScope<DevBuf<double>, DevBuf<double>> scope(d_x, d_o);
    SFB_CUDA(c, d_x.reserve(n));
    SFB_CUDA(c, d_o.reserve(n));
    SFB_CUDA(c, cudaMemcpyAsync(d_x.p, x, n * 8, cudaMemcpyHostToDevice, c->stream));
    k_digamma<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(d_x.p, n, d_o.p, exp_form);
    c->launches++;
    SFB_CUDA(c, cudaGetLastError());
    SFB_CUDA(c, cudaMemcpyAsync(out, d_o.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    SFB_CUDA(c, cudaStreamSynchronize(c->stream));
    return SFB200_OK;
}
