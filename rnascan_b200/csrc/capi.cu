// Library-level pieces of the C ABI: error reporting, device info, buffer sizing.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void rs_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int rs_cuda_fail(cudaError_t e, const char *what)
{
    rs_set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return RS_ERR_CUDA;
}

int rs_current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
    return dev < RS_MAX_DEVICES ? dev : RS_MAX_DEVICES - 1;
}

int rs_sm_count()
{
    static int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached = v;
        cached_dev = dev;
    }
    return cached;
}

// SMs the persistent scan kernels leave free (rs_set_reserved_sms): a collective that runs BESIDE a scan
// (the all-reduce of the background counts, device.BackgroundFusedScan) needs somewhere to be scheduled --
// a persistent kernel that fills every SM's shared memory would make it wait for the scan to end.
static thread_local int g_reserved_sms = 0;      // per calling thread: set -> launch -> reset never races another caller
extern "C" int rs_set_reserved_sms(int n)
{
    if (n < 0 || n > 64) { rs_set_error("rs_set_reserved_sms: 0..64"); return RS_ERR_INVALID; }
    g_reserved_sms = n;
    return RS_OK;
}
int rs_grid_sms()
{
    const int sms = rs_sm_count() - g_reserved_sms;
    return sms > 1 ? sms : 1;
}

// --------------------------------------------------------------------------- profiling hook
// A ring of CUDA event pairs recorded tightly around the main scan kernel of each entry
// point, on the caller's stream, so bench.py can report that kernel's own duration inside
// a longer timed step without synchronising between steps.
#include <vector>
static thread_local std::vector<cudaEvent_t> g_prof_ev;     // 2 * max_records; per calling thread
static thread_local int g_prof_n = 0, g_prof_cap = 0;
static thread_local bool g_prof_on = false, g_prof_open = false;

void rs_prof_start(cudaStream_t s)
{
    if (!g_prof_on || g_prof_n >= g_prof_cap) return;
    cudaEventRecord(g_prof_ev[2 * g_prof_n], s);
    g_prof_open = true;
}
void rs_prof_stop(cudaStream_t s)
{
    if (!g_prof_on || !g_prof_open) return;
    cudaEventRecord(g_prof_ev[2 * g_prof_n + 1], s);
    g_prof_open = false;
    g_prof_n++;
}

extern "C" int rs_prof_begin(int max_records)
{
    if (max_records < 1 || max_records > 65536) { rs_set_error("rs_prof_begin: bad max_records"); return RS_ERR_INVALID; }
    while ((int)g_prof_ev.size() < 2 * max_records) {
        cudaEvent_t e;
        RS_CUDA(cudaEventCreate(&e));
        g_prof_ev.push_back(e);
    }
    g_prof_cap = max_records; g_prof_n = 0; g_prof_on = true; g_prof_open = false;
    return RS_OK;
}

extern "C" int rs_prof_end(float *ms_out, int capacity, int *n_records)
{
    g_prof_on = false;
    int n = g_prof_n < capacity ? g_prof_n : capacity;
    for (int k = 0; k < n; k++) {
        RS_CUDA(cudaEventSynchronize(g_prof_ev[2 * k + 1]));
        RS_CUDA(cudaEventElapsedTime(&ms_out[k], g_prof_ev[2 * k], g_prof_ev[2 * k + 1]));
    }
    if (n_records) *n_records = n;
    return RS_OK;
}

extern "C" int rs_version(void) { return 100; }            // 0.1.0

extern "C" const char *rs_last_error(void) { return g_err; }

extern "C" int rs_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0, n = 0;
    RS_CUDA(cudaGetDeviceCount(&n));
    if (n <= 0) { rs_set_error("no CUDA device"); return RS_ERR_CUDA; }
    RS_CUDA(cudaGetDevice(&dev));
    int sm = 0, ma = 0, mi = 0;
    RS_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    RS_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
    RS_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sm;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    return RS_OK;
}

extern "C" int64_t rs_padded_count(int64_t n)
{
    if (n < 0) n = 0;
    return rs_roundup(n, RS_PAD) + RS_PAD;
}

// workspace: staging hit arrays (capacity entries) + per-tile segment table + scan scratch
WorkLayout rs_work_layout(int64_t n, int64_t capacity)
{
    WorkLayout wl;
    const int64_t cap = capacity > 0 ? capacity : 0;
    const int64_t tiles = (n > 0 ? n : 0) / RS_MIN_TILE + 16;
    int64_t off = 0;
    wl.off_pos = off;  off += rs_roundup(cap * 8, 256);
    wl.off_str = off;  off += rs_roundup(cap * 8, 256);
    wl.off_seq = off;  off += rs_roundup(cap * 4, 256);
    wl.off_seg = off;  off += rs_roundup(tiles * 16, 256);
    wl.off_scan = off; off += rs_roundup(rs_order_tmp_bytes(tiles), 256);
    wl.off_lut = off;  off += rs_kmer_work_bytes(n);      // k-mer scan: table, hit masks, segment scan
    wl.total = off;
    return wl;
}

extern "C" int64_t rs_scan_workspace_bytes(int64_t n, int64_t hit_capacity)
{
    return rs_work_layout(n, hit_capacity).total;
}
