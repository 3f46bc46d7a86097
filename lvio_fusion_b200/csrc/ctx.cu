// ctx.cu -- context, error text, launch accounting and the NCCL hook of liblvio_b200.so.
//
// NCCL is resolved with dlopen at run time (libnccl.so.2: the copy torch already mapped into
// the process if there is one, else the system library), so the library carries no link-time
// dependency on it and single-GPU use never touches it.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "lvb_internal.cuh"

namespace lvb {

static thread_local std::string g_last_error;
thread_local cudaStream_t g_alloc_stream = nullptr;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_last_error = buf;
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return LVB_ERR_CUDA;
}

// ---- NCCL through dlopen ---------------------------------------------------------------
struct NcclId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclId, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_comm_destroy)(void*);
typedef const char* (*fn_get_error_string)(int);

static struct {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_get_error_string get_error_string = nullptr;
} g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return LVB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { set_error("NCCL not found: %s", dlerror()); return LVB_ERR_COMM; }
    g_nccl.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
    g_nccl.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
    g_nccl.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.get_error_string = (fn_get_error_string)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce) { set_error("NCCL symbols missing"); return LVB_ERR_COMM; }
    g_nccl.handle = h;
    return LVB_OK;
}
static int nccl_fail(int rc, const char* what) {
    set_error("NCCL error %d (%s) in %s", rc, g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "?", what);
    return LVB_ERR_COMM;
}

int comm_allreduce_sum_f64(lvb_ctx* ctx, double* buf, size_t count) {
    if (ctx->world <= 1 || !ctx->comm) return LVB_OK;
    const int rc = g_nccl.all_reduce(buf, buf, count, /*ncclDouble*/ 8, /*ncclSum*/ 0, ctx->comm, ctx->stream);
    if (rc != 0) return nccl_fail(rc, "ncclAllReduce");
    return LVB_OK;
}

int comm_allreduce_min_i32(lvb_ctx* ctx, int* buf, size_t count) {
    if (ctx->world <= 1 || count == 0) return LVB_OK;
    const int rc = g_nccl.all_reduce(buf, buf, count, /*ncclInt32*/ 2, /*ncclMin*/ 3, ctx->comm, ctx->stream);
    if (rc != 0) return nccl_fail(rc, "ncclAllReduce(min)");
    return LVB_OK;
}

}  // namespace lvb

using namespace lvb;

extern "C" {

int lvb_version(void) { return 100; }

void lvb_default_options(lvb_solve_options* o) {
    o->max_num_iterations = 50; o->max_solver_time_in_seconds = 1e9; o->function_tolerance = 1e-6;
    o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8; o->initial_trust_region_radius = 1e4;
    o->jacobi_scaling = 1; o->linear_solver_type = 0; o->num_threads = 1; o->schur_mode = 0;
}

const char* lvb_last_error(void) { return g_last_error.c_str(); }

int lvb_ctx_create(int device, void* cuda_stream, lvb_ctx** out) {
    if (!out) { set_error("out is NULL"); return LVB_ERR_INVALID; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return LVB_ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range (%d devices)", device, count); return LVB_ERR_INVALID; }
    LVB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LVB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return LVB_ERR_CUDA; }
    lvb_ctx* c = new lvb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cuda_stream) { c->stream = (cudaStream_t)cuda_stream; c->own_stream = false; }
    else { LVB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    {   // keep freed device memory in the stream-ordered pool (see DevBuf)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { unsigned long long thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr); }
    }
    LVB_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    LVB_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    LVB_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    const char* no_side = getenv("LVB_NO_SIDE");
    c->use_side = !(no_side && no_side[0] == '1');
    const char* no_tma = getenv("LVB_NO_TMA");
    c->use_tma = !(no_tma && no_tma[0] == '1');
    const char* no_graph = getenv("LVB_NO_GRAPH");
    c->use_graph = !(no_graph && no_graph[0] == '1');
    const char* ev = getenv("LVB_EVAL_VARIANT");
    if (ev) c->eval_variant = atoi(ev);
    const char* ce = getenv("LVB_CHECK_EVERY");
    if (ce && atoi(ce) > 0) c->check_every = atoi(ce);
    *out = c;
    return LVB_OK;
}

void lvb_ctx_destroy(lvb_ctx* ctx) {
    if (!ctx) return;
    if (ctx->comm && g_nccl.comm_destroy) g_nccl.comm_destroy(ctx->comm);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

long long lvb_launch_count(lvb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int lvb_ctx_synchronize(lvb_ctx* ctx) {
    if (!ctx) return LVB_ERR_INVALID;
    LVB_CUDA(cudaStreamSynchronize(ctx->stream));
    return LVB_OK;
}

int lvb_comm_unique_id(char id[128]) {
    LVB_TRY(nccl_load());
    NcclId u;
    const int rc = g_nccl.get_unique_id(&u);
    if (rc != 0) return nccl_fail(rc, "ncclGetUniqueId");
    memcpy(id, u.internal, 128);
    return LVB_OK;
}

int lvb_comm_init(lvb_ctx* ctx, int rank, int world_size, const char id[128]) {
    if (!ctx || world_size < 1 || rank < 0 || rank >= world_size) { set_error("bad comm arguments"); return LVB_ERR_INVALID; }
    if (world_size == 1) { ctx->rank = 0; ctx->world = 1; return LVB_OK; }
    LVB_TRY(nccl_load());
    LVB_CUDA(cudaSetDevice(ctx->device));
    NcclId u; memcpy(u.internal, id, 128);
    void* comm = nullptr;
    const int rc = g_nccl.comm_init_rank(&comm, world_size, u, rank);
    if (rc != 0) return nccl_fail(rc, "ncclCommInitRank");
    ctx->comm = comm; ctx->rank = rank; ctx->world = world_size;
    return LVB_OK;
}

}  // extern "C"
