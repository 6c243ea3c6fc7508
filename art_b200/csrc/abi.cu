// libart_hotpath.so -- C-ABI entry points, context, host<->device staging.
// The compute lives in rcd.cu / amaze.cu / ...; this file holds no pixel arithmetic.
#include "ctx.h"

#include <functional>

#include <algorithm>
#include <thread>
#include <utility>

namespace {

constexpr size_t kStageBytes = 32u << 20;   // per half; two halves double-buffer pageable copies

int host_threads()
{
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 4;
    return (int)std::min(n, 16u);
}

// copy `nrows` rows of `wbytes` bytes between a row-pointer table and a packed buffer, in parallel
void rows_memcpy(const float* const* rows, int row0, int nrows, size_t wbytes, char* packed, bool to_packed)
{
    const int nt = std::max(1, std::min(host_threads(), nrows / 64));
    auto work = [&](int t) {
        const int a = (int)((long long)nrows * t / nt), b = (int)((long long)nrows * (t + 1) / nt);
        for (int i = a; i < b; ++i) {
            if (to_packed) memcpy(packed + (size_t)i * wbytes, rows[row0 + i], wbytes);
            else memcpy(const_cast<float*>(rows[row0 + i]), packed + (size_t)i * wbytes, wbytes);
        }
    };
    if (nt == 1) { work(0); return; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}

// true when rows[i] = rows[0] + i*stride for one stride (what array2D / PlanarPtr allocate)
bool constant_stride(const float* const* rows, int H, ptrdiff_t* stride)
{
    if (H < 2) { *stride = 0; return true; }
    const ptrdiff_t s = rows[1] - rows[0];
    for (int i = 2; i < H; ++i)
        if (rows[i] - rows[i - 1] != s) return false;
    *stride = s;
    return s > 0;
}

bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int ensure_stage(art_hp_ctx* ctx)
{
    if (ctx->h_stage[0]) return ART_HP_OK;
    for (int i = 0; i < 2; ++i) ART_CUDA(ctx, cudaMallocHost(&ctx->h_stage[i], kStageBytes));
    ctx->h_stage_bytes = kStageBytes;
    return ART_HP_OK;
}

struct Plane {
    const float* const* rows;   // host row table
    float* dev;                 // device plane
};

// Move planes between host row tables and device planes (pitch in floats) on `st`.
// Pinned, constant-stride host memory goes straight through cudaMemcpy2DAsync; anything else is
// staged through the context's two pinned halves, host memcpy overlapping the DMA of the other half.
int transfer(art_hp_ctx* ctx, cudaStream_t st, const Plane* planes, int nplanes, int W, int row_begin, int row_end, size_t pitch, bool to_device)
{
    const int H = row_end - row_begin;
    if (H <= 0) return ART_HP_OK;
    const size_t wbytes = (size_t)W * sizeof(float);
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    bool staged_any = false;
    int half = 0;
    bool busy[2] = {false, false};
    // for D2H we must copy out of a half only after its DMA finished; remember what to drain
    struct Pending { int plane, row0, nrows; } pend[2] = {{0, 0, 0}, {0, 0, 0}};
    auto drain = [&](int h) -> int {
        if (!busy[h]) return ART_HP_OK;
        ART_CUDA(ctx, cudaEventSynchronize(ctx->ev[h]));
        if (!to_device)
            rows_memcpy(planes[pend[h].plane].rows + row_begin, pend[h].row0, pend[h].nrows, wbytes, (char*)ctx->h_stage[h], false);
        busy[h] = false;
        return ART_HP_OK;
    };
    for (int p = 0; p < nplanes; ++p) {
        ptrdiff_t stride = 0;
        const float* const* hrows = planes[p].rows + row_begin;
        float* dbase = planes[p].dev + (size_t)row_begin * pitch;
        const bool cs = constant_stride(hrows, H, &stride);
        if (cs && is_pinned(hrows[0])) {
            const size_t spitch = (H > 1 ? (size_t)stride : (size_t)W) * sizeof(float);
            if (to_device)
                ART_CUDA(ctx, cudaMemcpy2DAsync(dbase, pitch * sizeof(float), hrows[0], spitch, wbytes, H, kind, st));
            else
                ART_CUDA(ctx, cudaMemcpy2DAsync(const_cast<float*>(hrows[0]), spitch, dbase, pitch * sizeof(float), wbytes, H, kind, st));
            continue;
        }
        staged_any = true;
        int rc = ensure_stage(ctx);
        if (rc) return rc;
        const int chunk = (int)std::max<size_t>(1, ctx->h_stage_bytes / wbytes);
        for (int r0 = 0; r0 < H; r0 += chunk) {
            const int n = std::min(chunk, H - r0);
            rc = drain(half);
            if (rc) return rc;
            char* hs = (char*)ctx->h_stage[half];
            float* d = dbase + (size_t)r0 * pitch;
            if (to_device) {
                rows_memcpy(hrows, r0, n, wbytes, hs, true);
                ART_CUDA(ctx, cudaMemcpy2DAsync(d, pitch * sizeof(float), hs, wbytes, wbytes, n, kind, st));
            } else {
                ART_CUDA(ctx, cudaMemcpy2DAsync(hs, wbytes, d, pitch * sizeof(float), wbytes, n, kind, st));
                pend[half] = {p, r0, n};
            }
            ART_CUDA(ctx, cudaEventRecord(ctx->ev[half], st));
            busy[half] = true;
            half ^= 1;
        }
    }
    if (staged_any) {
        int rc = drain(half);
        if (rc) return rc;
        rc = drain(half ^ 1);
        if (rc) return rc;
    }
    return ART_HP_OK;
}

bool rgb_bayer(unsigned filters)
{
    // RawImage::FC (rtengine/rawimage.h L186-189) over the 8x2 period of the descriptor: rows must
    // repeat with period 2, no site may be colour 3 (rcd_demosaic.cc L56-66), and the 2x2 cell must be
    // G on one diagonal with R and B on the other.
    auto FC = [filters](int r, int c) { return (filters >> ((((r) << 1 & 14) + ((c) & 1)) << 1)) & 3u; };
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 2; ++c)
            if (FC(r, c) == 3u || FC(r, c) != FC(r & 1, c)) return false;
    const unsigned f00 = FC(0, 0), f01 = FC(0, 1), f10 = FC(1, 0), f11 = FC(1, 1);
    return (f00 == 1 && f11 == 1 && f01 != 1 && f10 != 1 && f01 != f10) ||
           (f01 == 1 && f10 == 1 && f00 != 1 && f11 != 1 && f00 != f11);
}

}  // namespace

int art_reserve(art_hp_ctx* ctx, DevBuf& b, size_t bytes)
{
    if (b.bytes >= bytes) return ART_HP_OK;
    if (b.p) {
        ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ART_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.bytes = 0;
    }
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(ART_HP_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    b.bytes = bytes;
    return ART_HP_OK;
}

int art_pool_alloc(art_hp_ctx* ctx, size_t bytes, void** out)
{
    PoolBlk* best = nullptr;
    for (PoolBlk& b : ctx->pool)
        if (!b.used && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
    if (best) { best->used = true; *out = best->p; return ART_HP_OK; }
    // no fit: drop the free blocks that are too small (they would only pile up), then allocate
    for (size_t i = 0; i < ctx->pool.size();) {
        if (!ctx->pool[i].used) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->pool[i].p);
            ctx->pool.erase(ctx->pool.begin() + i);
        } else ++i;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(ART_HP_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    ctx->pool.push_back(PoolBlk{p, bytes, true});
    *out = p;
    return ART_HP_OK;
}

void art_pool_free(art_hp_ctx* ctx, void* p)
{
    for (PoolBlk& b : ctx->pool)
        if (b.p == p) { b.used = false; return; }
}

extern "C" {

int art_hp_abi_version(void) { return ART_HP_ABI_VERSION; }

int art_hp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int art_hp_create(art_hp_ctx** out, int device_id)
{
    if (!out) return ART_HP_ERR_INVALID;
    *out = nullptr;
    int n = art_hp_device_count();
    if (n <= 0) return ART_HP_ERR_NO_DEVICE;
    if (device_id < 0 || device_id >= n) return ART_HP_ERR_INVALID;
    if (cudaSetDevice(device_id) != cudaSuccess) { cudaGetLastError(); return ART_HP_ERR_CUDA; }
    art_hp_ctx* ctx = new art_hp_ctx();
    ctx->device = device_id;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 4; ++i) ok = cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); art_hp_destroy(ctx); return ART_HP_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return ART_HP_OK;
}

void art_hp_destroy(art_hp_ctx* ctx)
{
    if (ctx) art_hp_comm_destroy(ctx);
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->d_raw, &ctx->d_out[0], &ctx->d_out[1], &ctx->d_out[2], &ctx->d_scratch, &ctx->d_small, &ctx->d_work, &ctx->d_dn, &ctx->d_fattal, &ctx->d_small2, &ctx->d_dm[0], &ctx->d_dm[1], &ctx->d_dm[2], &ctx->d_chain_stages};
    if (ctx->d_nlm_dbg) cudaFree(ctx->d_nlm_dbg);
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    if (ctx->d_dn_tables.p) cudaFree(ctx->d_dn_tables.p);
    if (ctx->d_dn_labtabs.p) cudaFree(ctx->d_dn_labtabs.p);
    if (ctx->d_chain.p) cudaFree(ctx->d_chain.p);
    if (ctx->d_usm_tables.p) cudaFree(ctx->d_usm_tables.p);
    if (ctx->d_bl_lut.p) cudaFree(ctx->d_bl_lut.p);
    if (ctx->d_xt_cbrt.p) cudaFree(ctx->d_xt_cbrt.p);
    for (int i = 0; i < art_hp_ctx::NLANES; ++i) { if (ctx->lane[i]) cudaStreamDestroy(ctx->lane[i]); if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (auto& q : ctx->q) {
        if (q.raw.p) cudaFree(q.raw.p);
        for (DevBuf& o : q.out) if (o.p) cudaFree(o.p);
        if (q.packed.p) cudaFree(q.packed.p);
        if (q.up) cudaEventDestroy(q.up);
        if (q.done) cudaEventDestroy(q.done);
        if (q.down) cudaEventDestroy(q.down);
    }
    if (ctx->h_chain) cudaFreeHost(ctx->h_chain);
    if (ctx->ev_chain) cudaEventDestroy(ctx->ev_chain);
    for (PoolBlk& b : ctx->pool) cudaFree(b.p);
    for (int i = 0; i < 2; ++i) if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    delete ctx;
}

const char* art_hp_last_error(const art_hp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int art_hp_set_stream(art_hp_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return ART_HP_OK;
}

void* art_hp_get_stream(art_hp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int art_hp_sync(art_hp_ctx* ctx)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

unsigned long long art_hp_launch_count(const art_hp_ctx* ctx) { return ctx ? ctx->launches : 0ull; }

int art_hp_profile_enable(art_hp_ctx* ctx, int on)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& sp : ctx->spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
    ctx->spans.clear();
    ctx->stats.clear();
    ctx->profiling = on != 0;
    return ART_HP_OK;
}

int art_hp_profile_collect(art_hp_ctx* ctx)
{
    if (!ctx) return -1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    for (auto& sp : ctx->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
        ProfStat* st = nullptr;
        for (auto& x : ctx->stats) if (x.name == sp.name) { st = &x; break; }
        if (!st) { ctx->stats.push_back(ProfStat{sp.name, 0.0, 0}); st = &ctx->stats.back(); }
        st->ms += ms;
        st->calls += 1;
        cudaEventDestroy(sp.e0);
        cudaEventDestroy(sp.e1);
    }
    ctx->spans.clear();
    return (int)ctx->stats.size();
}

int art_hp_profile_entry(art_hp_ctx* ctx, int index, const char** name, double* total_ms, int* calls)
{
    if (!ctx || index < 0 || index >= (int)ctx->stats.size()) return ART_HP_ERR_INVALID;
    if (name) *name = ctx->stats[index].name.c_str();
    if (total_ms) *total_ms = ctx->stats[index].ms;
    if (calls) *calls = ctx->stats[index].calls;
    return ART_HP_OK;
}

void* art_hp_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void art_hp_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int art_hp_band_align(int method, int* period, int* offset)
{
    int p, o;
    switch (method) {
    case ART_HP_BAYER_AMAZE: p = 128; o = 0; break;     // tiles at stride ts-32 from -16, writing [16, ts-16)
    case ART_HP_BAYER_RCD: p = 176; o = 9; break;       // tiles at stride 176, writing [9, 185)
    default: return ART_HP_ERR_UNSUPPORTED;
    }
    if (period) *period = p;
    if (offset) *offset = o;
    return ART_HP_OK;
}

int art_hp_band_halo(int method)
{
    switch (method) {
    case ART_HP_BAYER_AMAZE: return 16;
    case ART_HP_BAYER_RCD: return 9;
    default: return -1;
    }
}

int art_hp_demosaic_bayer_rows_dev(art_hp_ctx* ctx, int method, int W, int H, unsigned filters,
                                   const float* d_raw, size_t raw_pitch,
                                   float* d_red, float* d_green, float* d_blue, size_t out_pitch,
                                   double initialGain, int border, int row_begin, int row_end)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw || !d_red || !d_green || !d_blue) return ctx->fail(ART_HP_ERR_INVALID, "null plane pointer");
    if (W < 32 || H < 32 || W > 65536 || H > 65536) return ctx->fail(ART_HP_ERR_INVALID, "frame %dx%d out of range [32,65536]", W, H);
    if (raw_pitch < (size_t)W || out_pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "pitch smaller than width");
    if (!rgb_bayer(filters)) return ctx->fail(ART_HP_ERR_INVALID, "filters=0x%08x is not an RGB Bayer pattern", filters);
    int P = 0, O = 0;
    if (art_hp_band_align(method, &P, &O)) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "unknown bayer method %d", method);
    auto aligned = [&](int r, int edge) { return r == edge || (r > O && r < H && (r - O) % P == 0); };
    if (row_begin < 0 || row_end > H || row_begin >= row_end || !aligned(row_begin, 0) || !aligned(row_end, H))
        return ctx->fail(ART_HP_ERR_INVALID, "rows [%d,%d) are not cut on the tile grid (period %d, offset %d) of a %d-row frame", row_begin, row_end, P, O, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    switch (method) {
    case ART_HP_BAYER_RCD:
        return art_rcd_dev(ctx, W, H, filters, d_raw, raw_pitch, d_red, d_green, d_blue, out_pitch, row_begin, row_end);
    default:
        if (!(initialGain > 0.0)) return ctx->fail(ART_HP_ERR_INVALID, "initialGain must be > 0");
        return art_amaze_dev(ctx, W, H, filters, d_raw, raw_pitch, d_red, d_green, d_blue, out_pitch, initialGain, border, row_begin, row_end);
    }
}

int art_hp_demosaic_bayer_dev(art_hp_ctx* ctx, int method, int W, int H, unsigned filters,
                              const float* d_raw, size_t raw_pitch,
                              float* d_red, float* d_green, float* d_blue, size_t out_pitch,
                              double initialGain, int border)
{
    return art_hp_demosaic_bayer_rows_dev(ctx, method, W, H, filters, d_raw, raw_pitch, d_red, d_green, d_blue, out_pitch,
                                          initialGain, border, 0, H);
}

int art_hp_border_interpolate2_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, int lborders,
                                   const float* d_raw, size_t raw_pitch,
                                   float* d_red, float* d_green, float* d_blue, size_t out_pitch)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw || !d_red || !d_green || !d_blue) return ctx->fail(ART_HP_ERR_INVALID, "null plane pointer");
    if (lborders < 1 || 2 * lborders >= W || 2 * lborders >= H) return ctx->fail(ART_HP_ERR_INVALID, "border %d does not fit %dx%d", lborders, W, H);
    if (!rgb_bayer(filters)) return ctx->fail(ART_HP_ERR_INVALID, "filters=0x%08x is not an RGB Bayer pattern", filters);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_border_dev(ctx, W, H, filters, lborders, d_raw, raw_pitch, d_red, d_green, d_blue, out_pitch, 0, H);
}

int art_hp_demosaic_bayer(art_hp_ctx* ctx, int method, int W, int H, unsigned filters,
                          const float* const* rawData,
                          float* const* red, float* const* green, float* const* blue,
                          double initialGain, int border)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!rawData || !red || !green || !blue) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    if (W < 32 || H < 32 || W > 65536 || H > 65536) return ctx->fail(ART_HP_ERR_INVALID, "frame %dx%d out of range [32,65536]", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_raw, plane))) return rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane in = {rawData, (float*)ctx->d_raw.p};
    Plane out[3] = {{red, (float*)ctx->d_out[0].p}, {green, (float*)ctx->d_out[1].p}, {blue, (float*)ctx->d_out[2].p}};
    if (method == ART_HP_BAYER_AMAZE && border >= 4 && initialGain > 0.0 && rgb_bayer(filters)) {
        // Banded pipeline: the raw rows of band k+1 travel host->device and the finished rows of band k-1
        // travel device->host (two DMA engines) while band k is being demosaiced.  Bands are whole
        // reference tile rows, so results do not change.
        struct Cb {
            art_hp_ctx* ctx; const Plane* in; const Plane* out; int W; size_t pitch; int uploaded;
            std::vector<cudaEvent_t> evs; std::vector<std::pair<int, int>> rows; std::vector<cudaEvent_t> upl;
        } cb{ctx, &in, out, W, pitch, 0, {}, {}, {}};
        auto hook = [](void* u, int phase, int r0, int r1) -> int {
            Cb* c = (Cb*)u;
            cudaEvent_t e;
            if (phase == 0) {
                if (r1 <= c->uploaded) return ART_HP_OK;
                int rc2 = transfer(c->ctx, c->ctx->copy_stream, c->in, 1, c->W, c->uploaded, r1, c->pitch, true);
                if (rc2) return rc2;
                c->uploaded = r1;
                if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return c->ctx->fail(ART_HP_ERR_CUDA, "cudaEventCreate failed");
                c->upl.push_back(e);
                cudaEventRecord(e, c->ctx->copy_stream);
                if (cudaStreamWaitEvent(c->ctx->stream, e, 0) != cudaSuccess) return c->ctx->fail(ART_HP_ERR_CUDA, "cudaStreamWaitEvent failed");
                return ART_HP_OK;
            }
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return c->ctx->fail(ART_HP_ERR_CUDA, "cudaEventCreate failed");
            cudaEventRecord(e, c->ctx->stream);
            c->evs.push_back(e);
            c->rows.push_back({r0, r1});
            return ART_HP_OK;
        };
        const int nty = (H + 16 + 127) / 128;
        const int band = std::max(2, (nty + 5) / 6);          // ~6 bands per frame
        rc = art_amaze_dev_banded(ctx, W, H, filters, (const float*)ctx->d_raw.p, pitch,
                                  (float*)ctx->d_out[0].p, (float*)ctx->d_out[1].p, (float*)ctx->d_out[2].p, pitch,
                                  initialGain, border, band, hook, &cb, 0, H);
        // all kernels are queued; now drain band by band on the device->host stream
        for (size_t k = 0; k < cb.evs.size() && rc == ART_HP_OK; ++k) {
            if (cudaStreamWaitEvent(ctx->d2h_stream, cb.evs[k], 0) != cudaSuccess) { rc = ctx->fail(ART_HP_ERR_CUDA, "cudaStreamWaitEvent failed"); break; }
            rc = transfer(ctx, ctx->d2h_stream, out, 3, W, cb.rows[k].first, cb.rows[k].second, pitch, false);
        }
        cudaError_t e1 = cudaStreamSynchronize(ctx->d2h_stream), e2 = cudaStreamSynchronize(ctx->stream);
        cudaError_t e0 = cudaStreamSynchronize(ctx->copy_stream);
        for (cudaEvent_t e : cb.evs) cudaEventDestroy(e);
        for (cudaEvent_t e : cb.upl) cudaEventDestroy(e);
        if (rc) return rc;
        if (e0 != cudaSuccess) return ctx->fail(ART_HP_ERR_CUDA, "copy stream sync failed: %s", cudaGetErrorString(e0));
        if (e1 != cudaSuccess || e2 != cudaSuccess) return ctx->fail(ART_HP_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        return ART_HP_OK;
    }
    if ((rc = transfer(ctx, ctx->stream, &in, 1, W, 0, H, pitch, true))) return rc;
    rc = art_hp_demosaic_bayer_dev(ctx, method, W, H, filters, (const float*)ctx->d_raw.p, pitch,
                                   (float*)ctx->d_out[0].p, (float*)ctx->d_out[1].p, (float*)ctx->d_out[2].p, pitch,
                                   initialGain, border);
    if (rc) return rc;
    if ((rc = transfer(ctx, ctx->stream, out, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_boxblur_dev(art_hp_ctx* ctx, const float* d_src, size_t src_pitch, float* d_dst, size_t dst_pitch, int radius, int W, int H)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_src || !d_dst) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || radius < 0 || src_pitch < (size_t)W || dst_pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d radius %d", W, H, radius);
    if (radius > 0 && (2 * radius + 1 > W || 2 * radius + 1 > H)) return ctx->fail(ART_HP_ERR_INVALID, "radius %d does not fit %dx%d", radius, W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_boxblur_dev(ctx, d_src, src_pitch, d_dst, dst_pitch, W, H, radius);
}

int art_hp_boxblur(art_hp_ctx* ctx, float* const* src, float* const* dst, int radius, int W, int H)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!src || !dst) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    if (W < 1 || H < 1 || radius < 0) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d radius %d", W, H, radius);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32), plane = pitch * (size_t)H * sizeof(float);
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_out[0], plane))) return rc;
    if ((rc = art_reserve(ctx, ctx->d_out[1], plane))) return rc;
    Plane in = {src, (float*)ctx->d_out[0].p};
    if ((rc = transfer(ctx, ctx->stream, &in, 1, W, 0, H, pitch, true))) return rc;
    float* dd = (src == dst) ? in.dev : (float*)ctx->d_out[1].p;
    if ((rc = art_hp_boxblur_dev(ctx, in.dev, pitch, dd, pitch, radius, W, H))) return rc;
    Plane out = {dst, dd};
    if ((rc = transfer(ctx, ctx->stream, &out, 1, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_guided_filter_dev(art_hp_ctx* ctx, int W, int H, const float* d_guide, size_t guide_pitch,
                             const float* d_src, size_t src_pitch, float* d_dst, size_t dst_pitch,
                             int r, float epsilon, int subsampling)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_guide || !d_src || !d_dst) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 8 || H < 8 || r < 1 || guide_pitch < (size_t)W || src_pitch < (size_t)W || dst_pitch < (size_t)W)
        return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d r %d", W, H, r);
    if (subsampling > 0 && (W / subsampling < 4 || H / subsampling < 4)) return ctx->fail(ART_HP_ERR_INVALID, "subsampling %d too coarse", subsampling);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_guided_dev(ctx, d_guide, guide_pitch, d_src, src_pitch, d_dst, dst_pitch, W, H, r, epsilon, subsampling);
}

int art_hp_guided_filter(art_hp_ctx* ctx, int W, int H, float* const* guide, float* const* src, float* const* dst,
                         int r, float epsilon, int subsampling)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!guide || !src || !dst) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    if (W < 8 || H < 8 || r < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d r %d", W, H, r);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32), plane = pitch * (size_t)H * sizeof(float);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane in[2] = {{guide, (float*)ctx->d_out[0].p}, {src, (float*)ctx->d_out[1].p}};
    const bool same = (guide == src);
    if ((rc = transfer(ctx, ctx->stream, in, same ? 1 : 2, W, 0, H, pitch, true))) return rc;
    const float* ds = same ? in[0].dev : in[1].dev;
    float* dd = (float*)ctx->d_out[2].p;
    if ((rc = art_hp_guided_filter_dev(ctx, W, H, in[0].dev, pitch, ds, pitch, dd, pitch, r, epsilon, subsampling))) return rc;
    Plane out = {dst, dd};
    if ((rc = transfer(ctx, ctx->stream, &out, 1, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_gauss_dev(art_hp_ctx* ctx, const float* d_src, size_t src_pitch, float* d_dst, size_t dst_pitch,
                     int W, int H, double sigma, int gausstype)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_src || !d_dst) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (gausstype != 0) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "only GAUSS_STANDARD is on the hot path");
    if (W < 4 || H < 4 || src_pitch < (size_t)W || dst_pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (!(sigma >= 0.0)) return ctx->fail(ART_HP_ERR_INVALID, "sigma must be >= 0");
    if (d_src == d_dst && src_pitch != dst_pitch) return ctx->fail(ART_HP_ERR_INVALID, "in-place call with two pitches");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_gauss_dev(ctx, d_src, src_pitch, d_dst, dst_pitch, W, H, sigma);
}

int art_hp_gauss(art_hp_ctx* ctx, float* const* src, float* const* dst, int W, int H, double sigma, int gausstype)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!src || !dst) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    if (gausstype != 0) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "only GAUSS_STANDARD is on the hot path");
    if (W < 4 || H < 4) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    const bool inplace = (src == dst);                       // the reference compares the row tables (gauss.cc L1441, L1447)
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_out[0], plane))) return rc;
    if (!inplace && (rc = art_reserve(ctx, ctx->d_out[1], plane))) return rc;
    float* ds = (float*)ctx->d_out[0].p;
    float* dd = inplace ? ds : (float*)ctx->d_out[1].p;
    Plane in = {src, ds};
    if ((rc = transfer(ctx, ctx->stream, &in, 1, W, 0, H, pitch, true))) return rc;
    if ((rc = art_gauss_dev(ctx, ds, pitch, dd, pitch, W, H, sigma))) return rc;
    Plane out = {dst, dd};
    if ((rc = transfer(ctx, ctx->stream, &out, 1, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

static int check_denoise_params(art_hp_ctx* ctx, const art_hp_denoise_params* P, const double* wprof, bool allow_auto = false)
{
    if (!P || !wprof) return ctx->fail(ART_HP_ERR_INVALID, "null parameters");
    if (P->colorSpace != 0 && P->colorSpace != 1) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "colorSpace %d", P->colorSpace);
    if (P->chrominanceMethod != 0 && !(allow_auto && P->chrominanceMethod == 1))
        return ctx->fail(ART_HP_ERR_UNSUPPORTED, "chrominanceMethod %d: AUTOMATIC needs the camera-space frame (art_hp_develop / art_hp_denoise_compute_params)", P->chrominanceMethod);
    if (P->colorSpace == 1 && !P->wprof_inverse) return ctx->fail(ART_HP_ERR_INVALID, "colorSpace LAB needs wprof_inverse");
    if (!(P->scale > 0) || !(P->gamma > 0)) return ctx->fail(ART_HP_ERR_INVALID, "scale and gamma must be positive");
    return ART_HP_OK;
}

int art_hp_rgb_denoise_dev(art_hp_ctx* ctx, float* d_r, float* d_g, float* d_b, size_t pitch, int W, int H,
                           const art_hp_denoise_params* params, const double wprof[9],
                           const float* d_cl_r, const float* d_cl_g, const float* d_cl_b, size_t cl_pitch, float* nresi_highresi)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b) return ctx->fail(ART_HP_ERR_INVALID, "null plane");
    if (W < 1 || H < 1 || W > 32767 || H > 32767 || pitch < (size_t)W)        // `short int imheight, imwidth`, FTblockDN.cc L1779
        return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d (sides are short ints in the reference)", W, H);
    int rc = check_denoise_params(ctx, params, wprof);
    if (rc) return rc;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_rgb_denoise_dev(ctx, d_r, d_g, d_b, pitch, W, H, params, wprof, d_cl_r, d_cl_g, d_cl_b, cl_pitch, nresi_highresi);
}

int art_hp_rgb_denoise(art_hp_ctx* ctx, float* const* r, float* const* g, float* const* b, int W, int H,
                       const art_hp_denoise_params* params, const double wprof[9],
                       float* const* cl_r, float* const* cl_g, float* const* cl_b, float* nresi_highresi)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    if (W < 1 || H < 1 || W > 32767 || H > 32767) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    int rc = check_denoise_params(ctx, params, wprof);
    if (rc) return rc;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    for (int c = 0; c < 3; ++c) if ((rc = art_reserve(ctx, ctx->d_out[c], plane))) return rc;
    float* d[3] = {(float*)ctx->d_out[0].p, (float*)ctx->d_out[1].p, (float*)ctx->d_out[2].p};
    Plane io[3] = {{r, d[0]}, {g, d[1]}, {b, d[2]}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    const int w2 = (W + 1) / 2, h2 = (H + 1) / 2;
    const size_t cp = round_up((size_t)w2, 32);
    float* dc[3] = {nullptr, nullptr, nullptr};
    if (cl_r && cl_g && cl_b) {
        if ((rc = art_reserve(ctx, ctx->d_raw, 3 * cp * (size_t)h2 * sizeof(float)))) return rc;
        for (int c = 0; c < 3; ++c) dc[c] = (float*)ctx->d_raw.p + (size_t)c * cp * h2;
        Plane cl[3] = {{cl_r, dc[0]}, {cl_g, dc[1]}, {cl_b, dc[2]}};
        if ((rc = transfer(ctx, ctx->stream, cl, 3, w2, 0, h2, cp, true))) return rc;
    }
    if ((rc = art_rgb_denoise_dev(ctx, d[0], d[1], d[2], pitch, W, H, params, wprof, dc[0], dc[1], dc[2], cp, nresi_highresi))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_detail_mask_dev(art_hp_ctx* ctx, const float* d_src, size_t src_pitch, float* d_mask, size_t mask_pitch, int W, int H,
                           float scaling, float threshold, float ceiling, float factor, int blur_type, float blur)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_src || !d_mask) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || src_pitch < (size_t)W || mask_pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (blur_type < 0 || blur_type > 2) return ctx->fail(ART_HP_ERR_INVALID, "blur_type %d (0 off, 1 box, 2 gauss)", blur_type);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = art_reserve(ctx, ctx->d_work, 2 * (size_t)(W / 4 + 1) * (H / 4 + 1) * sizeof(float));
    if (rc) return rc;
    return art_detail_mask_dev(ctx, d_src, src_pitch, d_mask, mask_pitch, W, H, scaling, threshold, ceiling, factor, blur_type, blur,
                               (float*)ctx->d_work.p);
}

int art_hp_detail_mask(art_hp_ctx* ctx, float* const* src, float* const* mask, int W, int H,
                       float scaling, float threshold, float ceiling, float factor, int blur_type, float blur)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!src || !mask) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_out[0], plane))) return rc;
    if ((rc = art_reserve(ctx, ctx->d_out[1], plane))) return rc;
    Plane in = {src, (float*)ctx->d_out[0].p};
    if ((rc = transfer(ctx, ctx->stream, &in, 1, W, 0, H, pitch, true))) return rc;
    if ((rc = art_hp_detail_mask_dev(ctx, (float*)ctx->d_out[0].p, pitch, (float*)ctx->d_out[1].p, pitch, W, H, scaling, threshold, ceiling,
                                     factor, blur_type, blur))) return rc;
    Plane out = {mask, (float*)ctx->d_out[1].p};
    if ((rc = transfer(ctx, ctx->stream, &out, 1, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_nlmeans_dev(art_hp_ctx* ctx, float* d_img, size_t pitch, int W, int H, float normcoeff, int strength, int detail_thresh, float scale)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_img) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_nlmeans_dev(ctx, d_img, pitch, W, H, normcoeff, strength, detail_thresh, scale);
}

int art_hp_nlmeans(art_hp_ctx* ctx, float* const* img, int W, int H, float normcoeff, int strength, int detail_thresh, float scale)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!img) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (!strength) return ART_HP_OK;                         // nlmeans.cc L52-54
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_out[0], pitch * (size_t)H * sizeof(float)))) return rc;
    float* d = (float*)ctx->d_out[0].p;
    Plane io = {img, d};
    if ((rc = transfer(ctx, ctx->stream, &io, 1, W, 0, H, pitch, true))) return rc;
    if ((rc = art_nlmeans_dev(ctx, d, pitch, W, H, normcoeff, strength, detail_thresh, scale))) return rc;
    if ((rc = transfer(ctx, ctx->stream, &io, 1, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_scale_colors_bayer_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* d_raw, size_t pitch,
                                  const float cblacksom[4], const float scale_mul[4], float chmax[3])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw || !cblacksom || !scale_mul || !chmax) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 2 || H < 2 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d pitch %zu", W, H, pitch);
    if (!rgb_bayer(filters)) return ctx->fail(ART_HP_ERR_INVALID, "filters=0x%08x is not an RGB Bayer pattern", filters);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = art_reserve(ctx, ctx->d_small, 256);
    if (rc) return rc;
    if ((rc = art_scale_colors_dev(ctx, W, H, filters, d_raw, pitch, cblacksom, scale_mul, (int*)ctx->d_small.p))) return rc;
    int bits[3];
    ART_CUDA(ctx, cudaMemcpyAsync(bits, ctx->d_small.p, sizeof bits, cudaMemcpyDeviceToHost, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; ++i) memcpy(&chmax[i], &bits[i], sizeof(float));
    return ART_HP_OK;
}

int art_hp_scale_colors_bayer(art_hp_ctx* ctx, int W, int H, unsigned filters, float* const* rawData,
                              const float cblacksom[4], const float scale_mul[4], float chmax[3])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!rawData || !cblacksom || !scale_mul || !chmax) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 2 || H < 2) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_raw, pitch * (size_t)H * sizeof(float)))) return rc;
    Plane io = {rawData, (float*)ctx->d_raw.p};
    if ((rc = transfer(ctx, ctx->stream, &io, 1, W, 0, H, pitch, true))) return rc;
    if ((rc = art_hp_scale_colors_bayer_dev(ctx, W, H, filters, io.dev, pitch, cblacksom, scale_mul, chmax))) return rc;
    if ((rc = transfer(ctx, ctx->stream, &io, 1, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_scale_colors_xtrans_dev(art_hp_ctx* ctx, int W, int H, const int xtrans[36], float* d_raw, size_t pitch,
                                   const float cblacksom[3], const float scale_mul[3], float chmax[3])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!xtrans || !d_raw || !cblacksom || !scale_mul || !chmax) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d pitch %zu", W, H, pitch);
    for (int i = 0; i < 36; ++i)
        if (xtrans[i] < 0 || xtrans[i] > 2) return ctx->fail(ART_HP_ERR_INVALID, "xtrans[%d] = %d is not a colour", i, xtrans[i]);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = art_reserve(ctx, ctx->d_small, 256);
    if (rc) return rc;
    if ((rc = art_scale_colors_xtrans_dev(ctx, W, H, xtrans, d_raw, pitch, cblacksom, scale_mul, (int*)ctx->d_small.p))) return rc;
    int bits[3];
    ART_CUDA(ctx, cudaMemcpyAsync(bits, ctx->d_small.p, sizeof bits, cudaMemcpyDeviceToHost, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; ++i) memcpy(&chmax[i], &bits[i], sizeof(float));
    return ART_HP_OK;
}

int art_hp_scale_colors_xtrans(art_hp_ctx* ctx, int W, int H, const int xtrans[36], float* const* rawData,
                               const float cblacksom[3], const float scale_mul[3], float chmax[3])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!xtrans || !rawData || !cblacksom || !scale_mul || !chmax) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_raw, pitch * (size_t)H * sizeof(float)))) return rc;
    Plane io = {rawData, (float*)ctx->d_raw.p};
    if ((rc = transfer(ctx, ctx->stream, &io, 1, W, 0, H, pitch, true))) return rc;
    if ((rc = art_hp_scale_colors_xtrans_dev(ctx, W, H, xtrans, io.dev, pitch, cblacksom, scale_mul, chmax))) return rc;
    if ((rc = transfer(ctx, ctx->stream, &io, 1, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_green_equilibrate_global_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* d_raw, size_t pitch, int border)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 2 || H < 2 || pitch < (size_t)W || border < 0) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d pitch %zu border %d", W, H, pitch, border);
    if (!rgb_bayer(filters)) return ctx->fail(ART_HP_ERR_INVALID, "filters=0x%08x is not an RGB Bayer pattern", filters);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_green_equilibrate_global_dev(ctx, W, H, filters, d_raw, pitch, border);
}

int art_hp_green_equilibrate_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* d_raw, size_t pitch, float thresh,
                                 const float* d_thresh_map, size_t map_pitch)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 2 || H < 2 || pitch < (size_t)W || (d_thresh_map && map_pitch < (size_t)W)) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d pitch %zu", W, H, pitch);
    if (!rgb_bayer(filters)) return ctx->fail(ART_HP_ERR_INVALID, "filters=0x%08x is not an RGB Bayer pattern", filters);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_green_equilibrate_dev(ctx, W, H, filters, d_raw, pitch, thresh, d_thresh_map, map_pitch);
}

static int green_eq_host(art_hp_ctx* ctx, int W, int H, unsigned filters, float* const* rawData, int global, int border, float thresh, const float* const* thresh_map)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!rawData) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 2 || H < 2) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_raw, pitch * (size_t)H * sizeof(float)))) return rc;
    Plane io = {rawData, (float*)ctx->d_raw.p};
    if ((rc = transfer(ctx, ctx->stream, &io, 1, W, 0, H, pitch, true))) return rc;
    if (global) rc = art_hp_green_equilibrate_global_dev(ctx, W, H, filters, io.dev, pitch, border);
    else {
        float* dmap = nullptr;
        if (thresh_map) {
            if ((rc = art_reserve(ctx, ctx->d_out[0], pitch * (size_t)H * sizeof(float)))) return rc;
            Plane mp = {const_cast<float* const*>(thresh_map), (float*)ctx->d_out[0].p};
            if ((rc = transfer(ctx, ctx->stream, &mp, 1, W, 0, H, pitch, true))) return rc;
            dmap = mp.dev;
        }
        rc = art_hp_green_equilibrate_dev(ctx, W, H, filters, io.dev, pitch, thresh, dmap, pitch);
    }
    if (rc) return rc;
    if ((rc = transfer(ctx, ctx->stream, &io, 1, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}
int art_hp_green_equilibrate_global(art_hp_ctx* ctx, int W, int H, unsigned filters, float* const* rawData, int border)
{
    return green_eq_host(ctx, W, H, filters, rawData, 1, border, 0.f, nullptr);
}
int art_hp_green_equilibrate(art_hp_ctx* ctx, int W, int H, unsigned filters, float* const* rawData, float thresh, const float* const* thresh_map)
{
    return green_eq_host(ctx, W, H, filters, rawData, 0, 0, thresh, thresh_map);
}

int art_hp_resize_lanczos_dev(art_hp_ctx* ctx, int sW, int sH, const float* d_s0, const float* d_s1, const float* d_s2, size_t src_pitch,
                              int dW, int dH, float* d_d0, float* d_d1, float* d_d2, size_t dst_pitch, float scale)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_s0 || !d_s1 || !d_s2 || !d_d0 || !d_d1 || !d_d2) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (sW < 1 || sH < 1 || dW < 1 || dH < 1 || src_pitch < (size_t)sW || dst_pitch < (size_t)dW || !(scale > 0.f))
        return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d -> %dx%d, scale %g", sW, sH, dW, dH, (double)scale);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_lanczos_dev(ctx, d_s0, d_s1, d_s2, src_pitch, sW, sH, d_d0, d_d1, d_d2, dst_pitch, dW, dH, scale);
}

int art_hp_resize_lanczos(art_hp_ctx* ctx, int sW, int sH, float* const* s0, float* const* s1, float* const* s2,
                          int dW, int dH, float* const* d0, float* const* d1, float* const* d2, float scale)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!s0 || !s1 || !s2 || !d0 || !d1 || !d2) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (sW < 1 || sH < 1 || dW < 1 || dH < 1 || !(scale > 0.f))
        return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d -> %dx%d, scale %g", sW, sH, dW, dH, (double)scale);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sp = round_up((size_t)sW, 32), dp = round_up((size_t)dW, 32);
    int rc;
    for (int i = 0; i < 3; ++i) {
        if ((rc = art_reserve(ctx, ctx->d_dm[i], sp * (size_t)sH * sizeof(float)))) return rc;
        if ((rc = art_reserve(ctx, ctx->d_out[i], dp * (size_t)dH * sizeof(float)))) return rc;
    }
    Plane in[3] = {{s0, (float*)ctx->d_dm[0].p}, {s1, (float*)ctx->d_dm[1].p}, {s2, (float*)ctx->d_dm[2].p}};
    Plane out[3] = {{d0, (float*)ctx->d_out[0].p}, {d1, (float*)ctx->d_out[1].p}, {d2, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, in, 3, sW, 0, sH, sp, true))) return rc;
    if ((rc = art_lanczos_dev(ctx, in[0].dev, in[1].dev, in[2].dev, sp, sW, sH, out[0].dev, out[1].dev, out[2].dev, dp, dW, dH, scale))) return rc;
    if ((rc = transfer(ctx, ctx->stream, out, 3, dW, 0, dH, dp, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_scale_convert_dev(art_hp_ctx* ctx, int W, int H, float* d_red, float* d_green, float* d_blue, size_t pitch,
                             const float mul[3], int doClip, const double mat[9])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_red || !d_green || !d_blue || !mul) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d pitch %zu", W, H, pitch);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_scale_convert_dev(ctx, W, H, d_red, d_green, d_blue, pitch, mul, doClip, mat);
}

int art_hp_scale_convert(art_hp_ctx* ctx, int W, int H, float* const* red, float* const* green, float* const* blue,
                         const float mul[3], int doClip, const double mat[9])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!red || !green || !blue || !mul) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane io[3] = {{red, (float*)ctx->d_out[0].p}, {green, (float*)ctx->d_out[1].p}, {blue, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_scale_convert_dev(ctx, W, H, io[0].dev, io[1].dev, io[2].dev, pitch, mul, doClip, mat))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_fattal_fast_dim(int dim) { return dim > 0 ? art_fattal_fast_dim(dim) : 0; }

int art_hp_fattal_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch,
                      int threshold, int amount, int satcontrol, const double ws[9])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !ws) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 3 || H < 3 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d pitch %zu", W, H, pitch);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_fattal_dev(ctx, d_r, d_g, d_b, pitch, W, H, threshold, amount, satcontrol, ws);
}

int art_hp_fattal(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b,
                  int threshold, int amount, int satcontrol, const double ws[9])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !ws) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 3 || H < 3) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_fattal_dev(ctx, io[0].dev, io[1].dev, io[2].dev, pitch, W, H, threshold, amount, satcontrol, ws))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_denoise_compute_params_dev(art_hp_ctx* ctx, int W, int H, const float* d_r, const float* d_g, const float* d_b, size_t pitch,
                                      const float mul[3], int doClip, const double cam2work[9], const double wprof[9], double gamma, int aggressive,
                                      float out3[3], float* stats)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !mul || !wprof || !out3) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 256 || H < 256 || W > 32767 || H > 32767 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (!(gamma > 0)) return ctx->fail(ART_HP_ERR_INVALID, "gamma must be positive");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_denoise_auto_chroma_dev(ctx, d_r, d_g, d_b, pitch, W, H, mul, doClip, cam2work, wprof, gamma, aggressive, out3, stats);
}

int art_hp_denoise_compute_params(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b,
                                  const float mul[3], int doClip, const double cam2work[9], const double wprof[9], double gamma, int aggressive,
                                  float out3[3], float* stats)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !mul || !wprof || !out3) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 256 || H < 256 || W > 32767 || H > 32767) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (!(gamma > 0)) return ctx->fail(ART_HP_ERR_INVALID, "gamma must be positive");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    return art_denoise_auto_chroma_dev(ctx, io[0].dev, io[1].dev, io[2].dev, pitch, W, H, mul, doClip, cam2work, wprof, gamma, aggressive, out3, stats);
}

int art_hp_develop_size(const art_hp_develop_params* params, int W, int H, int* out_w, int* out_h, int* border)
{
    if (!params) return ART_HP_ERR_INVALID;
    art_dev_geo g;
    const int rc = art_develop_geometry2(params, W, H, &g);
    if (out_w) *out_w = g.Wo;
    if (out_h) *out_h = g.Ho;
    if (border) *border = g.bd;
    return rc;
}

int art_hp_denoise_guided_smoothing_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch,
                                        const double ws[9], int guidedChromaRadius, double scale)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !ws) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W || guidedChromaRadius < 0) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d or radius %d", W, H, guidedChromaRadius);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_guided_smoothing_dev(ctx, d_r, d_g, d_b, pitch, W, H, ws, guidedChromaRadius, scale);
}

int art_hp_denoise_guided_smoothing(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b,
                                    const double ws[9], int guidedChromaRadius, double scale)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !ws) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || guidedChromaRadius < 0) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d or radius %d", W, H, guidedChromaRadius);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_guided_smoothing_dev(ctx, io[0].dev, io[1].dev, io[2].dev, pitch, W, H, ws, guidedChromaRadius, scale))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_color_chain_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch,
                           const art_hp_chain_params* params)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !params) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_chain_dev(ctx, W, H, d_r, d_g, d_b, pitch, params);
}

int art_hp_color_chain(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b,
                       const art_hp_chain_params* params)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !params) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_chain_dev(ctx, W, H, io[0].dev, io[1].dev, io[2].dev, pitch, params))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

static int lab_hist_finish(art_hp_ctx* ctx, const unsigned* d_hist, unsigned* hist16)
{
    ART_CUDA(ctx, cudaMemcpyAsync(hist16, d_hist, 65536 * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_lab_histogram_dev(art_hp_ctx* ctx, int W, int H, const float* d_r, const float* d_g, const float* d_b, size_t pitch,
                             const art_hp_chain_params* params, unsigned hist16[65536])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !params || !hist16) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_small2, 65536 * sizeof(unsigned)))) return rc;
    if ((rc = art_chain_lab_hist_dev(ctx, W, H, d_r, d_g, d_b, pitch, params, (unsigned*)ctx->d_small2.p))) return rc;
    return lab_hist_finish(ctx, (const unsigned*)ctx->d_small2.p, hist16);
}

int art_hp_lab_histogram(art_hp_ctx* ctx, int W, int H, const float* const* r, const float* const* g, const float* const* b,
                         const art_hp_chain_params* params, unsigned hist16[65536])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !params || !hist16) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    if ((rc = art_reserve(ctx, ctx->d_small2, 65536 * sizeof(unsigned)))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p},
                   {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_chain_lab_hist_dev(ctx, W, H, io[0].dev, io[1].dev, io[2].dev, pitch, params, (unsigned*)ctx->d_small2.p))) return rc;
    return lab_hist_finish(ctx, (const unsigned*)ctx->d_small2.p, hist16);
}

int art_hp_sharpen_usm_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch,
                           const art_hp_sharpen_params* params, const double ws[9])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !params || !ws) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (params->contrast < 0 || params->radius < 0) return ctx->fail(ART_HP_ERR_INVALID, "negative contrast or radius");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_usm_dev(ctx, d_r, d_g, d_b, pitch, W, H, params, ws);
}

int art_hp_sharpen_usm(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b,
                       const art_hp_sharpen_params* params, const double ws[9])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !params || !ws) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (params->contrast < 0 || params->radius < 0) return ctx->fail(ART_HP_ERR_INVALID, "negative contrast or radius");
    if (params->amount < 1 || W < 8 || H < 8) return ART_HP_OK;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_usm_dev(ctx, io[0].dev, io[1].dev, io[2].dev, pitch, W, H, params, ws))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

static int xtrans_check(art_hp_ctx* ctx, int passes, int W, int H, const int* xtrans)
{
    if (passes != 1 && passes != 3) return ctx->fail(ART_HP_ERR_INVALID, "passes must be 1 or 3, got %d", passes);
    if (W < 23 || H < 23 || W > 16380 || H > 65536) return ctx->fail(ART_HP_ERR_INVALID, "frame %dx%d out of range", W, H);
    int n[3] = {0, 0, 0};
    for (int i = 0; i < 36; ++i) {
        if (xtrans[i] < 0 || xtrans[i] > 2) return ctx->fail(ART_HP_ERR_INVALID, "xtrans[%d] = %d is not a colour", i, xtrans[i]);
        n[xtrans[i]]++;
    }
    if (n[1] != 20 || n[0] != 8 || n[2] != 8) return ctx->fail(ART_HP_ERR_INVALID, "not an X-Trans matrix (%d R, %d G, %d B)", n[0], n[1], n[2]);
    return ART_HP_OK;
}

int art_hp_demosaic_xtrans_dev(art_hp_ctx* ctx, int passes, int useCieLab, int W, int H, const int xtrans[36], const float rgb_cam[12],
                               const float* d_raw, size_t raw_pitch, float* d_red, float* d_green, float* d_blue, size_t out_pitch)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!xtrans || !rgb_cam || !d_raw || !d_red || !d_green || !d_blue) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    int rc = xtrans_check(ctx, passes, W, H, xtrans);
    if (rc) return rc;
    if (raw_pitch < (size_t)W || out_pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "pitch smaller than the width");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_xtrans_dev(ctx, passes, useCieLab != 0, W, H, xtrans, rgb_cam, d_raw, raw_pitch, d_red, d_green, d_blue, out_pitch);
}

int art_hp_demosaic_xtrans(art_hp_ctx* ctx, int passes, int useCieLab, int W, int H, const int xtrans[36], const float rgb_cam[12],
                           float* const* rawData, float* const* red, float* const* green, float* const* blue)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!xtrans || !rgb_cam || !rawData || !red || !green || !blue) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    int rc = xtrans_check(ctx, passes, W, H, xtrans);
    if (rc) return rc;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    if ((rc = art_reserve(ctx, ctx->d_raw, plane))) return rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane in = {rawData, (float*)ctx->d_raw.p};
    Plane out[3] = {{red, (float*)ctx->d_out[0].p}, {green, (float*)ctx->d_out[1].p}, {blue, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, &in, 1, W, 0, H, pitch, true))) return rc;
    if ((rc = art_xtrans_dev(ctx, passes, useCieLab != 0, W, H, xtrans, rgb_cam, in.dev, pitch, out[0].dev, out[1].dev, out[2].dev, pitch))) return rc;
    if ((rc = transfer(ctx, ctx->stream, out, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

// ---- hot / dead pixel filter (badpixels.cu)
static int read_count(art_hp_ctx* ctx, const int* d_count, int* count)
{
    int n = 0;
    ART_CUDA(ctx, cudaMemcpyAsync(&n, d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (count) *count = n;
    return ART_HP_OK;
}

int art_hp_find_hot_dead_pixels_dev(art_hp_ctx* ctx, int W, int H, const int* xtrans, const float* d_raw, size_t raw_pitch, float thresh,
                                    int findHotPixels, int findDeadPixels, unsigned char* d_map, size_t map_pitch, int* count)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw || !d_map) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || H > 65535 || raw_pitch < (size_t)W || map_pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (xtrans)
        for (int i = 0; i < 36; ++i)
            if (xtrans[i] < 0 || xtrans[i] > 2) return ctx->fail(ART_HP_ERR_INVALID, "xtrans[%d] = %d is not a colour", i, xtrans[i]);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = art_reserve(ctx, ctx->d_small2, 256);
    if (rc) return rc;
    if ((rc = art_find_hot_dead_dev(ctx, W, H, xtrans, d_raw, raw_pitch, thresh, findHotPixels != 0, findDeadPixels != 0, d_map, map_pitch, (int*)ctx->d_small2.p))) return rc;
    return read_count(ctx, (const int*)ctx->d_small2.p, count);
}

int art_hp_interpolate_bad_pixels_bayer_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* d_raw, size_t raw_pitch,
                                            const unsigned char* d_map, size_t map_pitch, int* count)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw || !d_map) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || H > 65535 || raw_pitch < (size_t)W || map_pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = art_reserve(ctx, ctx->d_small2, 256);
    if (rc) return rc;
    if ((rc = art_interpolate_bad_bayer_dev(ctx, W, H, filters, d_raw, raw_pitch, d_map, map_pitch, (int*)ctx->d_small2.p))) return rc;
    return read_count(ctx, (const int*)ctx->d_small2.p, count);
}

int art_hp_interpolate_bad_pixels_xtrans_dev(art_hp_ctx* ctx, int W, int H, const int* xtrans, float* d_raw, size_t raw_pitch,
                                             const unsigned char* d_map, size_t map_pitch, int* count)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw || !d_map || !xtrans) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || H > 65535 || raw_pitch < (size_t)W || map_pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = art_reserve(ctx, ctx->d_small2, 256);
    if (rc) return rc;
    if ((rc = art_interpolate_bad_xtrans_dev(ctx, W, H, xtrans, d_raw, raw_pitch, d_map, map_pitch, (int*)ctx->d_small2.p))) return rc;
    return read_count(ctx, (const int*)ctx->d_small2.p, count);
}

// host forms: the raw plane and the byte map travel to the device and back (the map through d_out[0], one byte per pixel)
static int badpix_host(art_hp_ctx* ctx, int W, int H, const int* xtrans, unsigned filters, const float* const* rawData, float thresh, int hot, int dead,
                       unsigned char* map, size_t map_stride, int* count, bool interpolate)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!rawData || !map) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || H > 65535 || map_stride < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32), mp = round_up((size_t)W, 128);
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_raw, pitch * (size_t)H * sizeof(float)))) return rc;
    if ((rc = art_reserve(ctx, ctx->d_out[0], mp * (size_t)H))) return rc;
    Plane io = {rawData, (float*)ctx->d_raw.p};
    unsigned char* d_map = (unsigned char*)ctx->d_out[0].p;
    if ((rc = transfer(ctx, ctx->stream, &io, 1, W, 0, H, pitch, true))) return rc;
    ART_CUDA(ctx, cudaMemcpy2DAsync(d_map, mp, map, map_stride, (size_t)W, (size_t)H, cudaMemcpyHostToDevice, ctx->stream));
    if (interpolate) {
        if ((rc = xtrans ? art_hp_interpolate_bad_pixels_xtrans_dev(ctx, W, H, xtrans, io.dev, pitch, d_map, mp, count)
                         : art_hp_interpolate_bad_pixels_bayer_dev(ctx, W, H, filters, io.dev, pitch, d_map, mp, count))) return rc;
        if ((rc = transfer(ctx, ctx->stream, &io, 1, W, 0, H, pitch, false))) return rc;
    } else {
        if ((rc = art_hp_find_hot_dead_pixels_dev(ctx, W, H, xtrans, io.dev, pitch, thresh, hot, dead, d_map, mp, count))) return rc;
        ART_CUDA(ctx, cudaMemcpy2DAsync(map, map_stride, d_map, mp, (size_t)W, (size_t)H, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_find_hot_dead_pixels(art_hp_ctx* ctx, int W, int H, const int* xtrans, const float* const* rawData, float thresh,
                                int findHotPixels, int findDeadPixels, unsigned char* map, size_t map_stride, int* count)
{
    return badpix_host(ctx, W, H, xtrans, 0u, rawData, thresh, findHotPixels, findDeadPixels, map, map_stride, count, false);
}

int art_hp_interpolate_bad_pixels_bayer(art_hp_ctx* ctx, int W, int H, unsigned filters, float* const* rawData,
                                        const unsigned char* map, size_t map_stride, int* count)
{
    return badpix_host(ctx, W, H, nullptr, filters, rawData, 0.f, 0, 0, const_cast<unsigned char*>(map), map_stride, count, true);
}

int art_hp_interpolate_bad_pixels_xtrans(art_hp_ctx* ctx, int W, int H, const int* xtrans, float* const* rawData,
                                         const unsigned char* map, size_t map_stride, int* count)
{
    if (ctx && !xtrans) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    return badpix_host(ctx, W, H, xtrans, 0u, rawData, 0.f, 0, 0, const_cast<unsigned char*>(map), map_stride, count, true);
}

int art_hp_channel_mixer_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch, const float matrix[9])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !matrix) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_channel_mixer_dev(ctx, W, H, d_r, d_g, d_b, pitch, matrix);
}

int art_hp_channel_mixer(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b, const float matrix[9])
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !matrix) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], pitch * (size_t)H * sizeof(float)))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_channel_mixer_dev(ctx, W, H, io[0].dev, io[1].dev, io[2].dev, pitch, matrix))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

// ---- HSL equalizer (hsl.cu)
int art_hp_hsl_equalizer_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch, const art_hp_hsl_params* params)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !params) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_hsl_equalizer_dev(ctx, W, H, d_r, d_g, d_b, pitch, params);
}

int art_hp_hsl_equalizer(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b, const art_hp_hsl_params* params)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !params) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], pitch * (size_t)H * sizeof(float)))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_hsl_equalizer_dev(ctx, W, H, io[0].dev, io[1].dev, io[2].dev, pitch, params))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

// ---- proPhotoBlue (bw.cu)
int art_hp_prophoto_blue_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_prophoto_blue_dev(ctx, W, H, d_r, d_g, d_b, pitch);
}

int art_hp_prophoto_blue(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], pitch * (size_t)H * sizeof(float)))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_prophoto_blue_dev(ctx, W, H, io[0].dev, io[1].dev, io[2].dev, pitch))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

// ---- black and white (bw.cu)
int art_hp_black_and_white_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch, const art_hp_bw_params* params)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !params) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_bw_dev(ctx, W, H, d_r, d_g, d_b, pitch, params);
}

int art_hp_black_and_white(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b, const art_hp_bw_params* params)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !params) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], pitch * (size_t)H * sizeof(float)))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_bw_dev(ctx, W, H, io[0].dev, io[1].dev, io[2].dev, pitch, params))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

// ---- tone equalizer (toneeq.cu)
int art_hp_tone_equalizer_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch, const art_hp_toneeq_params* params)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !params) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_tone_equalizer_dev(ctx, W, H, d_r, d_g, d_b, pitch, params);
}

int art_hp_tone_equalizer(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b, const art_hp_toneeq_params* params)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !params) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], pitch * (size_t)H * sizeof(float)))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_tone_equalizer_dev(ctx, W, H, io[0].dev, io[1].dev, io[2].dev, pitch, params))) return rc;
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

// ---- dual demosaic (dual.cu)
static int vng4_check(art_hp_ctx* ctx, int W, int H, unsigned prefilters)
{
    if (W < 8 || H < 8 || W > 65536 || H > 65535) return ctx->fail(ART_HP_ERR_INVALID, "frame %dx%d out of range", W, H);
    const unsigned filters = prefilters & ~((prefilters & 0x55555555u) << 1);
    if (!rgb_bayer(filters)) return ctx->fail(ART_HP_ERR_INVALID, "prefilters=0x%08x does not collapse to an RGB Bayer pattern", prefilters);
    return ART_HP_OK;
}

int art_hp_demosaic_vng4_dev(art_hp_ctx* ctx, int W, int H, unsigned prefilters, const float* d_raw, size_t raw_pitch,
                             float* d_red, float* d_green, float* d_blue, size_t out_pitch)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw || !d_red || !d_green || !d_blue) return ctx->fail(ART_HP_ERR_INVALID, "null plane pointer");
    int rc = vng4_check(ctx, W, H, prefilters);
    if (rc) return rc;
    if (raw_pitch < (size_t)W || out_pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "pitch smaller than width");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_vng4_dev(ctx, W, H, prefilters, d_raw, raw_pitch, d_red, d_green, d_blue, out_pitch);
}

static int dual_check(art_hp_ctx* ctx, int W, int H, double contrast, int autoContrast)
{
    if (!(contrast >= 0.0)) return ctx->fail(ART_HP_ERR_INVALID, "contrast must be >= 0");
    if (autoContrast && (W < 80 || H < 80)) return ctx->fail(ART_HP_ERR_INVALID, "the automatic contrast threshold needs a frame of at least 80x80, got %dx%d", W, H);
    return ART_HP_OK;
}

int art_hp_dual_demosaic_bayer_dev(art_hp_ctx* ctx, int method, int second, int W, int H, unsigned filters, unsigned prefilters,
                                   const float* d_raw, size_t raw_pitch, float* d_red, float* d_green, float* d_blue, size_t out_pitch,
                                   double initialGain, int border, double contrast, int autoContrast, float* d_threshold_out)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (second != ART_HP_DUAL_BILINEAR && second != ART_HP_DUAL_VNG4) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "unknown flat-region demosaicer %d", second);
    int rc = dual_check(ctx, W, H, contrast, autoContrast);
    if (rc) return rc;
    if (second == ART_HP_DUAL_VNG4) {
        if ((rc = vng4_check(ctx, W, H, prefilters))) return rc;
        if ((prefilters & ~((prefilters & 0x55555555u) << 1)) != filters) return ctx->fail(ART_HP_ERR_INVALID, "prefilters=0x%08x is not the four-colour form of filters=0x%08x", prefilters, filters);
    }
    if ((rc = art_hp_demosaic_bayer_dev(ctx, method, W, H, filters, d_raw, raw_pitch, d_red, d_green, d_blue, out_pitch, initialGain, border))) return rc;
    if (contrast == 0.0 && !autoContrast) {       // dual_demosaic_RT.cc L43-71: the first demosaicer alone
        if (d_threshold_out) ART_CUDA(ctx, cudaMemsetAsync(d_threshold_out, 0, sizeof(float), ctx->stream));
        return ART_HP_OK;
    }
    return art_dual_blend_dev(ctx, second, W, H, second == ART_HP_DUAL_VNG4 ? prefilters : filters, nullptr, d_raw, raw_pitch,
                              d_red, d_green, d_blue, out_pitch, contrast, autoContrast, d_threshold_out);
}

int art_hp_dual_demosaic_xtrans_dev(art_hp_ctx* ctx, int passes, int useCieLab, int W, int H, const int xtrans[36], const float rgb_cam[12],
                                    const float* d_raw, size_t raw_pitch, float* d_red, float* d_green, float* d_blue, size_t out_pitch,
                                    double contrast, int autoContrast, float* d_threshold_out)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    int rc = dual_check(ctx, W, H, contrast, autoContrast);
    if (rc) return rc;
    if ((rc = art_hp_demosaic_xtrans_dev(ctx, passes, useCieLab, W, H, xtrans, rgb_cam, d_raw, raw_pitch, d_red, d_green, d_blue, out_pitch))) return rc;
    if (contrast == 0.0 && !autoContrast) {
        if (d_threshold_out) ART_CUDA(ctx, cudaMemsetAsync(d_threshold_out, 0, sizeof(float), ctx->stream));
        return ART_HP_OK;
    }
    return art_dual_blend_dev(ctx, 2, W, H, 0, xtrans, d_raw, raw_pitch, d_red, d_green, d_blue, out_pitch, contrast, autoContrast, d_threshold_out);
}

// host entries: upload the raw plane, run `stage` on the device planes, download the frame, read the threshold back
typedef std::function<int(float*, size_t, float*, float*, float*, int*)> DualStage;       // sets *blended when the blend step ran
static int dual_host(art_hp_ctx* ctx, int W, int H, const float* const* rawData, float* const* red, float* const* green, float* const* blue,
                     double* contrast, const DualStage& stage)
{
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_raw, plane))) return rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane in = {rawData, (float*)ctx->d_raw.p};
    Plane out[3] = {{red, (float*)ctx->d_out[0].p}, {green, (float*)ctx->d_out[1].p}, {blue, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, &in, 1, W, 0, H, pitch, true))) return rc;
    int blended = 0;
    if ((rc = stage(in.dev, pitch, out[0].dev, out[1].dev, out[2].dev, &blended))) return rc;
    if ((rc = transfer(ctx, ctx->stream, out, 3, W, 0, H, pitch, false))) return rc;
    float thr = 0.f;
    if (contrast && blended) ART_CUDA(ctx, cudaMemcpyAsync(&thr, art_dual_threshold_slot(ctx), sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (contrast) *contrast = thr * 100.f;           // dual_demosaic_RT.cc L112: contrast = contrastf * 100.f
    return ART_HP_OK;
}

int art_hp_demosaic_vng4(art_hp_ctx* ctx, int W, int H, unsigned prefilters, const float* const* rawData,
                         float* const* red, float* const* green, float* const* blue)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!rawData || !red || !green || !blue) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    int rc = vng4_check(ctx, W, H, prefilters);
    if (rc) return rc;
    return dual_host(ctx, W, H, rawData, red, green, blue, nullptr, [&](float* raw, size_t p, float* r, float* g, float* b, int*) {
        return art_vng4_dev(ctx, W, H, prefilters, raw, p, r, g, b, p);
    });
}

int art_hp_dual_demosaic_bayer(art_hp_ctx* ctx, int method, int second, int W, int H, unsigned filters, unsigned prefilters,
                               const float* const* rawData, float* const* red, float* const* green, float* const* blue,
                               double initialGain, int border, double* contrast, int autoContrast)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!rawData || !red || !green || !blue || !contrast) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 32 || H < 32 || W > 65536 || H > 65535) return ctx->fail(ART_HP_ERR_INVALID, "frame %dx%d out of range", W, H);
    const double c0 = *contrast;
    return dual_host(ctx, W, H, rawData, red, green, blue, contrast, [&](float* raw, size_t p, float* r, float* g, float* b, int* blended) {
        *blended = c0 != 0.0 || autoContrast;
        return art_hp_dual_demosaic_bayer_dev(ctx, method, second, W, H, filters, prefilters, raw, p, r, g, b, p, initialGain, border, c0, autoContrast, nullptr);
    });
}

int art_hp_dual_demosaic_xtrans(art_hp_ctx* ctx, int passes, int useCieLab, int W, int H, const int xtrans[36], const float rgb_cam[12],
                                const float* const* rawData, float* const* red, float* const* green, float* const* blue,
                                double* contrast, int autoContrast)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!xtrans || !rgb_cam || !rawData || !red || !green || !blue || !contrast) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    int rc = xtrans_check(ctx, passes, W, H, xtrans);
    if (rc) return rc;
    const double c0 = *contrast;
    return dual_host(ctx, W, H, rawData, red, green, blue, contrast, [&](float* raw, size_t p, float* r, float* g, float* b, int* blended) {
        *blended = c0 != 0.0 || autoContrast;
        return art_hp_dual_demosaic_xtrans_dev(ctx, passes, useCieLab, W, H, xtrans, rgb_cam, raw, p, r, g, b, p, c0, autoContrast, nullptr);
    });
}

int art_hp_median_denoise_dev(art_hp_ctx* ctx, const float* d_src, size_t src_pitch, float* d_dst, size_t dst_pitch, int W, int H,
                              int median_type, int use_upper, float upper_bound)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_src || !d_dst || d_src == d_dst) return ctx->fail(ART_HP_ERR_INVALID, "null or aliased planes (the device form is out of place)");
    if (W < 1 || H < 1 || src_pitch < (size_t)W || dst_pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (median_type < 0 || median_type > 5) return ctx->fail(ART_HP_ERR_INVALID, "median_type %d", median_type);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_median_dev(ctx, d_src, src_pitch, d_dst, dst_pitch, W, H, median_type, use_upper, upper_bound);
}

int art_hp_median_denoise(art_hp_ctx* ctx, float* const* src, float* const* dst, int W, int H, int median_type,
                          int use_upper, float upper_bound)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!src || !dst) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (median_type < 0 || median_type > 5) return ctx->fail(ART_HP_ERR_INVALID, "median_type %d", median_type);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    int rc;
    for (int i = 0; i < 2; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    Plane in = {src, (float*)ctx->d_out[0].p}, out = {dst, (float*)ctx->d_out[1].p};
    if ((rc = transfer(ctx, ctx->stream, &in, 1, W, 0, H, pitch, true))) return rc;
    if ((rc = art_median_dev(ctx, in.dev, pitch, out.dev, pitch, W, H, median_type, use_upper, upper_bound))) return rc;
    if ((rc = transfer(ctx, ctx->stream, &out, 1, W, 0, H, pitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_redft00_2d(art_hp_ctx* ctx, int n0, int n1, const float* in, float* out)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!in || !out || n0 < 3 || n1 < 3) return ctx->fail(ART_HP_ERR_INVALID, "bad arguments");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(float) * (size_t)n0 * n1;
    int rc;
    if ((rc = art_reserve(ctx, ctx->d_out[0], bytes))) return rc;
    if ((rc = art_reserve(ctx, ctx->d_out[1], bytes))) return rc;
    ART_CUDA(ctx, cudaMemcpyAsync(ctx->d_out[0].p, in, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = art_redft00_2d_dev(ctx, (const float*)ctx->d_out[0].p, (float*)ctx->d_out[1].p, n0, n1))) return rc;
    ART_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_out[1].p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

static int check_develop(art_hp_ctx* ctx, const art_hp_develop_params* p, int W, int H)
{
    if (!p) return ctx->fail(ART_HP_ERR_INVALID, "null parameters");
    if (p->method == ART_HP_XTRANS_3PASS || p->method == ART_HP_XTRANS_1PASS) {
        if (!p->xtrans || !p->rgb_cam) return ctx->fail(ART_HP_ERR_INVALID, "X-Trans methods need xtrans and rgb_cam");
        int rc = xtrans_check(ctx, p->method == ART_HP_XTRANS_3PASS ? 3 : 1, W, H, p->xtrans);
        if (rc) return rc;
    } else {
        if (p->method != ART_HP_BAYER_AMAZE && p->method != ART_HP_BAYER_RCD) return ctx->fail(ART_HP_ERR_INVALID, "unknown demosaic method %d", p->method);
        if (!rgb_bayer(p->filters)) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "filters 0x%08x is not an RGB Bayer pattern", p->filters);
    }
    if (W < 32 || H < 32 || W > 32767 || H > 32767) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    {
        if (p->tran < 0 || p->tran > 15) return ctx->fail(ART_HP_ERR_INVALID, "tran %d is not a combination of TR_R90 / R180 / R270, TR_VFLIP, TR_HFLIP", p->tran);
        art_dev_geo g;
        if (art_develop_geometry2(p, W, H, &g))
            return ctx->fail(ART_HP_ERR_INVALID, "the PreviewProps window (%d, %d, %d x %d, skip %d) leaves the developed frame", p->pp_x, p->pp_y, p->pp_width, p->pp_height, p->pp_skip);
        if (g.Wo < 24 || g.Ho < 24) return ctx->fail(ART_HP_ERR_INVALID, "frame %dx%d (developed %dx%d) is too small for a border of %d", W, H, g.Wo, g.Ho, g.bd);
        if (g.skip > 64) return ctx->fail(ART_HP_ERR_INVALID, "pp_skip %d", g.skip);
    }
    if (p->guidedChromaRadius < 0) return ctx->fail(ART_HP_ERR_INVALID, "guidedChromaRadius %d", p->guidedChromaRadius);
    if (p->tran < 0 || p->tran > 15) return ctx->fail(ART_HP_ERR_INVALID, "tran %d is not a combination of TR_R90 / R180 / R270, TR_VFLIP, TR_HFLIP", p->tran);
    if (p->hr_blend && !(p->hlmax[0] > 0 && p->hlmax[1] > 0 && p->hlmax[2] > 0)) return ctx->fail(ART_HP_ERR_INVALID, "hr_blend needs positive hlmax");
    if ((p->denoise || p->fattal_enabled) && !p->wprof) return ctx->fail(ART_HP_ERR_INVALID, "wprof is required by denoise and tone mapping");
    if (p->denoise) { int rc = check_denoise_params(ctx, p->denoise, p->wprof, true); if (rc) return rc; }
    return ART_HP_OK;
}

int art_hp_develop_dev(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, const float* d_raw, size_t raw_pitch,
                       float* d_r, float* d_g, float* d_b, size_t out_pitch)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw || !d_r || !d_g || !d_b) return ctx->fail(ART_HP_ERR_INVALID, "null plane");
    int rc = check_develop(ctx, params, W, H);
    if (rc) return rc;
    int bd, Wo, Ho;
    art_develop_geometry(params, W, H, &bd, &Wo, &Ho);
    if (raw_pitch < (size_t)W || out_pitch < (size_t)Wo) return ctx->fail(ART_HP_ERR_INVALID, "pitch smaller than the width");
    if (params->chain && ((out_pitch & 3) || ((reinterpret_cast<uintptr_t>(d_r) | reinterpret_cast<uintptr_t>(d_g) | reinterpret_cast<uintptr_t>(d_b)) & 15)))
        return ctx->fail(ART_HP_ERR_INVALID, "the colour chain needs 16-byte aligned output planes with a pitch that is a multiple of 4 floats");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_develop_dev(ctx, params, W, H, d_raw, raw_pitch, d_r, d_g, d_b, out_pitch);
}

int art_hp_band_plan_rows(const art_hp_develop_params* params, int W, int H, int own_begin, int own_end, int halo, art_hp_band_plan* plan)
{
    if (!params || !plan || W < 32 || H < 32) return ART_HP_ERR_INVALID;
    return art_band_plan(params, W, H, own_begin, own_end, halo, plan);
}

int art_hp_develop_band_dev(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, const float* d_raw, size_t raw_pitch,
                            float* d_r, float* d_g, float* d_b, size_t out_pitch, const art_hp_band_plan* plan)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_raw || !d_r || !d_g || !d_b || !plan) return ctx->fail(ART_HP_ERR_INVALID, "null plane or plan");
    int rc = check_develop(ctx, params, W, H);
    if (rc) return rc;
    int bd, Wo, Ho;
    art_develop_geometry(params, W, H, &bd, &Wo, &Ho);
    if (raw_pitch < (size_t)W || out_pitch < (size_t)Wo) return ctx->fail(ART_HP_ERR_INVALID, "pitch smaller than the width");
    if (params->chain && ((out_pitch & 3) || ((reinterpret_cast<uintptr_t>(d_r) | reinterpret_cast<uintptr_t>(d_g) | reinterpret_cast<uintptr_t>(d_b)) & 15)))
        return ctx->fail(ART_HP_ERR_INVALID, "the colour chain needs 16-byte aligned output planes with a pitch that is a multiple of 4 floats");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_develop_band_dev(ctx, params, W, H, d_raw, raw_pitch, d_r, d_g, d_b, out_pitch, plan);
}

int art_hp_develop(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, float* const* rawData,
                   float* const* r, float* const* g, float* const* b)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!rawData || !r || !g || !b) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    int rc = check_develop(ctx, params, W, H);
    if (rc) return rc;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    int bd, Wo, Ho;
    art_develop_geometry(params, W, H, &bd, &Wo, &Ho);
    const size_t pitch = round_up((size_t)W, 32), opitch = round_up((size_t)Wo, 32);
    if ((rc = art_reserve(ctx, ctx->d_raw, pitch * (size_t)H * sizeof(float)))) return rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], opitch * (size_t)Ho * sizeof(float)))) return rc;
    Plane in = {rawData, (float*)ctx->d_raw.p};
    Plane out[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, &in, 1, W, 0, H, pitch, true))) return rc;
    if ((rc = art_develop_dev(ctx, params, W, H, in.dev, pitch, out[0].dev, out[1].dev, out[2].dev, opitch))) return rc;
    if ((rc = transfer(ctx, ctx->stream, out, 3, Wo, 0, Ho, opitch, false))) return rc;
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

// Batch-queue form: the copies of frame k overlap the kernels of frames k-1 / k+1 (three streams, two frames in flight).
// packed_out != nullptr: the frame leaves as interleaved scanlines (pack.cu) instead of three float planes.
static int develop_submit(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, float* const* rawData,
                          float* const* r, float* const* g, float* const* b, int bps, int is_float, void* packed_out, size_t packed_stride)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!rawData || (!packed_out && (!r || !g || !b))) return ctx->fail(ART_HP_ERR_INVALID, "null row table");
    int rc = check_develop(ctx, params, W, H);
    if (rc) return rc;
    if (ctx->q_submitted - ctx->q_collected >= 2) return ctx->fail(ART_HP_ERR_INVALID, "two frames are already in flight: call art_hp_develop_wait first");
    int bd, Wo, Ho;
    art_develop_geometry(params, W, H, &bd, &Wo, &Ho);
    ptrdiff_t st[4] = {0, 0, 0, 0};
    float* const* tabs[4] = {rawData, r, g, b};
    for (int i = 0; i < (packed_out ? 1 : 4); ++i)
        if (!constant_stride(tabs[i], i ? Ho : H, &st[i]) || !is_pinned(tabs[i][0]))
            return ctx->fail(ART_HP_ERR_UNSUPPORTED, "the batch queue needs pinned, constant-stride planes (art_hp_host_alloc); use art_hp_develop otherwise");
    size_t row_bytes = 0;
    if (packed_out) {
        if (art_scanline_mode(bps, is_float) < 0) return ctx->fail(ART_HP_ERR_INVALID, "bps %d / isFloat %d is not a format of Imagefloat::getScanline", bps, is_float);
        row_bytes = (size_t)Wo * 3 * (bps / 8);
        if (packed_stride < row_bytes || !is_pinned(packed_out)) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "the packed output must be pinned with rows of at least %zu bytes", row_bytes);
    }
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    art_hp_ctx::QSlot& q = ctx->q[ctx->q_submitted & 1];
    const size_t pitch = round_up((size_t)W, 32), opitch = round_up((size_t)Wo, 32);
    if ((rc = art_reserve(ctx, q.raw, pitch * (size_t)H * sizeof(float)))) return rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, q.out[i], opitch * (size_t)Ho * sizeof(float)))) return rc;
    const size_t dpitch = round_up(row_bytes, 128);
    if (packed_out && (rc = art_reserve(ctx, q.packed, dpitch * (size_t)Ho))) return rc;
    if (!q.up) {
        ART_CUDA(ctx, cudaEventCreateWithFlags(&q.up, cudaEventDisableTiming));
        ART_CUDA(ctx, cudaEventCreateWithFlags(&q.done, cudaEventDisableTiming));
        ART_CUDA(ctx, cudaEventCreateWithFlags(&q.down, cudaEventDisableTiming));
    }
    const size_t wbytes = (size_t)W * sizeof(float);
    // the slot's previous frame (k-2) has been collected, so its planes are free; the upload runs beside frame k-1's kernels
    ART_CUDA(ctx, cudaMemcpy2DAsync(q.raw.p, pitch * sizeof(float), rawData[0], (H > 1 ? (size_t)st[0] : (size_t)W) * sizeof(float), wbytes, H,
                                    cudaMemcpyHostToDevice, ctx->copy_stream));
    ART_CUDA(ctx, cudaEventRecord(q.up, ctx->copy_stream));
    ART_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, q.up, 0));
    if ((rc = art_develop_dev(ctx, params, W, H, (const float*)q.raw.p, pitch, (float*)q.out[0].p, (float*)q.out[1].p, (float*)q.out[2].p, opitch))) return rc;
    if (packed_out && (rc = art_scanlines_dev(ctx, Wo, Ho, (const float*)q.out[0].p, (const float*)q.out[1].p, (const float*)q.out[2].p, opitch, bps, is_float,
                                              q.packed.p, dpitch))) return rc;
    ART_CUDA(ctx, cudaEventRecord(q.done, ctx->stream));
    ART_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, q.done, 0));
    if (packed_out) {
        ART_CUDA(ctx, cudaMemcpy2DAsync(packed_out, packed_stride, q.packed.p, dpitch, row_bytes, Ho, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    } else {
        for (int i = 0; i < 3; ++i)
            ART_CUDA(ctx, cudaMemcpy2DAsync(tabs[i + 1][0], (size_t)st[i + 1] * sizeof(float), q.out[i].p, opitch * sizeof(float), (size_t)Wo * sizeof(float), Ho,
                                            cudaMemcpyDeviceToHost, ctx->d2h_stream));
    }
    ART_CUDA(ctx, cudaEventRecord(q.down, ctx->d2h_stream));
    ctx->q_submitted++;
    return ART_HP_OK;
}

int art_hp_develop_submit(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, float* const* rawData,
                          float* const* r, float* const* g, float* const* b)
{
    return develop_submit(ctx, params, W, H, rawData, r, g, b, 0, 0, nullptr, 0);
}

int art_hp_develop_submit_packed(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, float* const* rawData,
                                 int bps, int isFloat, void* out, size_t out_stride_bytes)
{
    if (ctx && !out) return ctx->fail(ART_HP_ERR_INVALID, "null output");
    return develop_submit(ctx, params, W, H, rawData, nullptr, nullptr, nullptr, bps, isFloat, out, out_stride_bytes);
}

int art_hp_scanlines_dev(art_hp_ctx* ctx, int W, int H, const float* d_r, const float* d_g, const float* d_b, size_t pitch, int bps, int isFloat,
                         void* d_out, size_t out_stride_bytes)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!d_r || !d_g || !d_b || !d_out) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    return art_scanlines_dev(ctx, W, H, d_r, d_g, d_b, pitch, bps, isFloat, d_out, out_stride_bytes);
}

int art_hp_scanlines(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b, int bps, int isFloat,
                     void* out, size_t out_stride_bytes)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (!r || !g || !b || !out) return ctx->fail(ART_HP_ERR_INVALID, "null pointer");
    if (W < 1 || H < 1) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (art_scanline_mode(bps, isFloat) < 0) return ctx->fail(ART_HP_ERR_INVALID, "bps %d / isFloat %d is not a format of Imagefloat::getScanline", bps, isFloat);
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pitch = round_up((size_t)W, 32);
    const size_t plane = pitch * (size_t)H * sizeof(float);
    const size_t row_bytes = (size_t)W * 3 * (bps / 8), dpitch = round_up(row_bytes, 128);
    if (out_stride_bytes < row_bytes) return ctx->fail(ART_HP_ERR_INVALID, "output rows need %zu bytes", row_bytes);
    int rc;
    for (int i = 0; i < 3; ++i)
        if ((rc = art_reserve(ctx, ctx->d_out[i], plane))) return rc;
    if ((rc = art_reserve(ctx, ctx->d_raw, dpitch * (size_t)H))) return rc;
    Plane io[3] = {{r, (float*)ctx->d_out[0].p}, {g, (float*)ctx->d_out[1].p}, {b, (float*)ctx->d_out[2].p}};
    if ((rc = transfer(ctx, ctx->stream, io, 3, W, 0, H, pitch, true))) return rc;
    if ((rc = art_scanlines_dev(ctx, W, H, io[0].dev, io[1].dev, io[2].dev, pitch, bps, isFloat, ctx->d_raw.p, dpitch))) return rc;
    ART_CUDA(ctx, cudaMemcpy2DAsync(out, out_stride_bytes, ctx->d_raw.p, dpitch, row_bytes, H, cudaMemcpyDeviceToHost, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_develop_wait(art_hp_ctx* ctx)
{
    if (!ctx) return ART_HP_ERR_INVALID;
    if (ctx->q_submitted == ctx->q_collected) return ctx->fail(ART_HP_ERR_INVALID, "no frame in flight");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    ART_CUDA(ctx, cudaEventSynchronize(ctx->q[ctx->q_collected & 1].down));
    ctx->q_collected++;
    return ART_HP_OK;
}

int art_hp_develop_pending(const art_hp_ctx* ctx) { return ctx ? (int)(ctx->q_submitted - ctx->q_collected) : 0; }

}  // extern "C"
