// vtk.cu -- writeOutputData (postprocess.h:5-47) with the text produced on the GPU.
//
// The reference prints N^3 coordinate lines ("%10.8e %10.8e %10.8e\n") and N^3 value lines
// ("%10.8e\n") with one fprintf each: ~2.5 minutes for the 8 GB of a 513^3 grid, six times
// its own CPU solve.  Here the file body is produced chunk by chunk on the device and
// streamed to the host through two pinned buffers (mgb_vtk_open / _next / _close): the
// caller only fwrite()s.  The bytes are the reference's, always:
//
//   * a coordinate line is three 14-character strings out of per-axis tables (h*i for
//     i < n: 3n values, formatted once on the HOST with snprintf -- h*i is a short dyadic
//     fraction whose tenth digit is often an exact 5, i.e. a rounding tie) plus separators:
//     fixed 45 bytes per point, pure byte shuffling on the device;
//   * a value line needs the correctly rounded 9-digit decimal of a double.  fmt_e8 computes
//     v * 10^(8-E) as a 53 x 128 -> 181-bit integer product with the 128-bit TRUNCATED power
//     of ten from pow10_table.h, so the 9 digits are exact and the discarded fraction is known
//     to within 2^-98 of a unit: the rounding is certain unless the fraction lies within that
//     distance below one half or looks like an exact tie through an inexact power.  Those
//     cases, non-finite values and three-digit exponents are EXCEPTIONS: the chunk that holds
//     one is formatted again on the host with snprintf("%10.8e\n") -- the C library decides,
//     never a guess.  (True ties through an exact power, 10^0..10^38, go to even on the
//     device, as printf does in round-to-nearest.)  tests/test_vtk_format_cpu.py runs the
//     same integer steps in Python against printf semantics.
//   * value lines are 15 or 16 bytes (sign): a block-wise count of the negatives and one
//     small scan give every block its output offset.
//
// Three kernels per value chunk (convert+count, scan, emit) and one per coordinate chunk;
// HBM-trivial -- the stream is bound by PCIe (8 GB at ~55 GB/s) and by the caller's
// fwrite, not by formatting any more.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mgb.h"
#include "pow10_table.h"

namespace mgb {

int set_error(const char *msg);  // api.cu: what mgb_last_error() returns
long long *launch_counter();     // kernels.cu

static int fail(const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return set_error(buf);
}

namespace {

constexpr int kFmtThreads = 256;
constexpr unsigned long long kExcMark = ~0ull;

// ---- the decimal conversion ------------------------------------------------------------
// packed result: bit 63 sign, bits 40..47 = E + 128, bits 0..39 = the 9 digits as an
// integer (10^8 <= D < 10^9, or 0 for +-0); kExcMark: exception
__device__ __forceinline__ unsigned long long fmt_e8(unsigned long long bits,
                                                     const uint64_t (*__restrict__ p10)[2],
                                                     const int16_t *__restrict__ p10e)
{
    const unsigned long long sign = bits & 0x8000000000000000ull;
    const int ex = (int)((bits >> 52) & 0x7FF);
    unsigned long long m = bits & 0xFFFFFFFFFFFFFull;
    if (ex == 0x7FF)
        return kExcMark;
    int e2;
    if (ex == 0) {
        if (m == 0)
            return sign | (128ull << 40);
        const int sh = __clzll((long long)m) - 11;  // bring the top bit to position 52
        m <<= sh;
        e2 = -1074 - sh;
    } else {
        m |= 1ull << 52;
        e2 = ex - 1075;
    }
    int E = ((e2 + 52) * 1233) >> 12;  // floor(log10(2^(e2+52))) or one less
    for (int attempt = 0; attempt < 3; attempt++) {
        const int q = 8 - E;
        if (q < kPow10Min || q > kPow10Max)
            return kExcMark;
        const unsigned long long ph = p10[q - kPow10Min][0], pl = p10[q - kPow10Min][1];
        const int s = -(e2 + (int)p10e[q - kPow10Min]);
        if (s < 130 || s > 190)
            return kExcMark;
        // m * P = r2:r1:r0 (181 bits)
        const unsigned long long r0 = m * pl;
        const unsigned long long c1 = __umul64hi(m, pl);
        const unsigned long long lo2 = m * ph;
        unsigned long long r1 = c1 + lo2;
        const unsigned long long r2 = __umul64hi(m, ph) + (r1 < c1 ? 1ull : 0ull);
        unsigned long long D = r2 >> (s - 128);
        if (D >= 1000000000ull) {
            E++;
            continue;
        }
        if (D < 100000000ull) {
            E--;
            continue;
        }
        const unsigned long long f2 = r2 & ((1ull << (s - 128)) - 1), h2 = 1ull << (s - 129);
        const bool exact = q >= 0 && q <= 38;  // 10^q < 2^128: the entry is 10^q itself
        if (f2 > h2 || (f2 == h2 && (r1 | r0) != 0)) {
            D++;
        } else if (f2 == h2) {
            if (!exact)
                return kExcMark;  // looks like a tie, but the power was truncated
            D += D & 1;           // a true tie: to even
        } else if (f2 == h2 - 1 && r1 == ~0ull && !exact) {
            return kExcMark;  // within the truncation error below one half
        }
        if (D == 1000000000ull) {
            D = 100000000ull;
            E++;
        }
        if (E < -99 || E > 99)
            return kExcMark;
        return sign | ((unsigned long long)(E + 128) << 40) | D;
    }
    return kExcMark;
}

// the characters of one packed value, without the newline; returns the length (14 or 15)
__device__ __forceinline__ int fmt_put(unsigned long long pk, char *o)
{
    int n = 0;
    if (pk >> 63)
        o[n++] = '-';
    unsigned int D = (unsigned int)(pk & 0xFFFFFFFFFFull);
    const int E = (int)((pk >> 40) & 0xFF) - 128;
    char dig[9];
#pragma unroll
    for (int i = 8; i >= 0; i--) {
        dig[i] = (char)('0' + D % 10u);
        D /= 10u;
    }
    o[n++] = dig[0];
    o[n++] = '.';
#pragma unroll
    for (int i = 1; i < 9; i++)
        o[n++] = dig[i];
    o[n++] = 'e';
    const int a = E < 0 ? -E : E;
    o[n++] = E < 0 ? '-' : '+';
    o[n++] = (char)('0' + a / 10);
    o[n++] = (char)('0' + a % 10);
    return n;
}

// ---- value chunk, kernel 1: convert, count negatives per block, count exceptions ----
__global__ void __launch_bounds__(kFmtThreads)
k_vtk_convert(const double *__restrict__ v, long long n, unsigned long long *__restrict__ packed,
              unsigned int *__restrict__ block_neg, unsigned int *__restrict__ exceptions,
              const uint64_t (*__restrict__ p10)[2], const int16_t *__restrict__ p10e)
{
    const long long i = (long long)blockIdx.x * kFmtThreads + threadIdx.x;
    unsigned long long pk = 0;
    bool neg = false, exc = false;
    if (i < n) {
        pk = fmt_e8((unsigned long long)__double_as_longlong(v[i]), p10, p10e);
        exc = pk == kExcMark;
        neg = !exc && (pk >> 63);
        packed[i] = pk;
    }
    __shared__ unsigned int s_neg, s_exc;
    if (threadIdx.x == 0)
        s_neg = s_exc = 0;
    __syncthreads();
    const unsigned int bn = __popc(__ballot_sync(0xffffffffu, neg));
    const unsigned int be = __popc(__ballot_sync(0xffffffffu, exc));
    if ((threadIdx.x & 31) == 0) {
        if (bn) atomicAdd(&s_neg, bn);
        if (be) atomicAdd(&s_exc, be);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        block_neg[blockIdx.x] = s_neg;
        if (s_exc)
            atomicAdd(exceptions, s_exc);
    }
}

// ---- kernel 2: exclusive scan of the per-block negative counts (one block) ----
__global__ void __launch_bounds__(1024)
k_vtk_scan(unsigned int *__restrict__ block_neg, int nblocks, unsigned long long *__restrict__ total_bytes,
           long long n)
{
    __shared__ unsigned int part[1024];
    const int t = threadIdx.x;
    const int per = (nblocks + 1023) / 1024;
    const int lo = t * per, hi = min(lo + per, nblocks);
    unsigned int sum = 0;
    for (int b = lo; b < hi; b++)
        sum += block_neg[b];
    part[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele, inclusive
        const unsigned int x = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += x;
        __syncthreads();
    }
    unsigned int run = t ? part[t - 1] : 0;
    for (int b = lo; b < hi; b++) {
        const unsigned int c = block_neg[b];
        block_neg[b] = run;
        run += c;
    }
    if (t == 1023)
        *total_bytes = 15ull * (unsigned long long)n + part[1023];
}

// ---- kernel 3: the text.  A block builds its <= 256 lines in shared memory and copies
// them to its place in the output ----
__global__ void __launch_bounds__(kFmtThreads)
k_vtk_emit(const unsigned long long *__restrict__ packed, long long n,
           const unsigned int *__restrict__ block_neg_excl, char *__restrict__ out)
{
    __shared__ char text[kFmtThreads * 16];
    __shared__ unsigned int warp_neg[kFmtThreads / 32];
    const long long first = (long long)blockIdx.x * kFmtThreads;
    const long long i = first + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long pk = 0;
    bool neg = false;
    if (i < n) {
        pk = packed[i];
        neg = pk != kExcMark && (pk >> 63);
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, neg);
    if (lane == 0)
        warp_neg[w] = __popc(bal);
    __syncthreads();
    unsigned int before = __popc(bal & ((1u << lane) - 1));
    for (int x = 0; x < w; x++)
        before += warp_neg[x];
    unsigned int total = 0;
    for (int x = 0; x < kFmtThreads / 32; x++)
        total += warp_neg[x];
    if (i < n && pk != kExcMark) {
        char *o = text + 15 * threadIdx.x + before;
        const int len = fmt_put(pk, o);
        o[len] = '\n';
    }
    __syncthreads();
    const long long cnt = min((long long)kFmtThreads, n - first);
    const unsigned int bytes = 15u * (unsigned int)cnt + total;
    char *dst = out + 15 * first + block_neg_excl[blockIdx.x];
    for (unsigned int b = threadIdx.x; b < bytes; b += kFmtThreads)
        dst[b] = text[b];
}

// ---- coordinate chunk: 45 bytes per point out of the three axis tables ----
// tab: [ni + nj + nk][16] (14 characters used); 4 output bytes per thread
__global__ void __launch_bounds__(256)
k_vtk_points(const char *__restrict__ tab, int ni, int nj, int nk, long long first_point,
             long long nbytes, unsigned int *__restrict__ out)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * w >= nbytes)
        return;
    // position of this thread's first byte: point p = (i, j, k), character c of its line
    long long p = first_point + (4 * w) / 45;
    int c = (int)((4 * w) % 45);
    int k = (int)(p % nk);
    const long long r = p / nk;
    int j = (int)(r % nj), i = (int)(r / nj);
    unsigned int word = 0;
#pragma unroll
    for (int x = 0; x < 4; x++) {
        char ch = 0;
        if (4 * w + x < nbytes) {
            if (c == 44)
                ch = '\n';
            else if (c == 14 || c == 29)
                ch = ' ';
            else if (c < 14)
                ch = tab[i * 16 + c];
            else if (c < 29)
                ch = tab[(ni + j) * 16 + c - 15];
            else
                ch = tab[(ni + nj + k) * 16 + c - 30];
        }
        word |= (unsigned int)(unsigned char)ch << (8 * x);
        if (++c == 45) {  // next point, k fastest (postprocess.h:22-33)
            c = 0;
            if (++k == nk) {
                k = 0;
                if (++j == nj) {
                    j = 0;
                    i++;
                }
            }
        }
    }
    out[w] = word;
}

constexpr long long kChunkPoints = 1 << 20;  // multiple of 4: chunk offsets stay 4-byte aligned
constexpr size_t kHeadroom = 256;            // in front of the text: room for the section headers
constexpr size_t kBufBytes = kHeadroom + 45 * (size_t)kChunkPoints + 64;

}  // namespace
}  // namespace mgb

using namespace mgb;

struct mgb_vtk {
    const double *values = nullptr;
    int ni = 0, nj = 0, nk = 0;
    long long total = 0, nchunks = 0;
    long long next_issue = 0, next_return = 0;  // stream positions: 0..nchunks-1 points, then values
    cudaStream_t st = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr};
    char *h_buf[2] = {nullptr, nullptr};  // pinned
    char *d_text = nullptr;
    double *d_vals = nullptr;
    unsigned long long *d_packed = nullptr;
    unsigned int *d_block_neg = nullptr;
    unsigned int *d_exc = nullptr;
    unsigned long long *d_total = nullptr;
    char *d_tab = nullptr;
    uint64_t (*d_p10)[2] = nullptr;
    int16_t *d_p10e = nullptr;
    struct Info {
        unsigned long long bytes;
        unsigned int exceptions;
    };
    Info *h_info = nullptr;  // pinned, [2]
    std::string head0, head1;
    long long host_chunks = 0;
    int device = 0;
};

namespace {

#define VK(call)                                                                 \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess)                                                   \
            return fail("%s: %s", #call, cudaGetErrorString(e_));                \
    } while (0)

void vtk_free(mgb_vtk *w)
{
    if (!w)
        return;
    cudaSetDevice(w->device);
    if (w->st) cudaStreamSynchronize(w->st);
    for (int b = 0; b < 2; b++) {
        if (w->h_buf[b]) cudaFreeHost(w->h_buf[b]);
        if (w->done[b]) cudaEventDestroy(w->done[b]);
    }
    if (w->h_info) cudaFreeHost(w->h_info);
    cudaFree(w->d_text);
    cudaFree(w->d_vals);
    cudaFree(w->d_packed);
    cudaFree(w->d_block_neg);
    cudaFree(w->d_exc);
    cudaFree(w->d_total);
    cudaFree(w->d_tab);
    cudaFree(w->d_p10);
    cudaFree(w->d_p10e);
    if (w->st) cudaStreamDestroy(w->st);
    delete w;
}

// enqueue stream position `pos` into buffer pos & 1
int vtk_issue(mgb_vtk *w, long long pos)
{
    const int b = (int)(pos & 1);
    const bool points = pos < w->nchunks;
    const long long c = points ? pos : pos - w->nchunks;
    const long long first = c * kChunkPoints;
    const long long cnt = std::min(kChunkPoints, w->total - first);
    char *dst = w->h_buf[b] + kHeadroom;
    if (points) {
        const long long nbytes = 45 * cnt;
        const long long words = (nbytes + 3) / 4;
        k_vtk_points<<<(unsigned)((words + 255) / 256), 256, 0, w->st>>>(
            w->d_tab, w->ni, w->nj, w->nk, first, nbytes, reinterpret_cast<unsigned int *>(w->d_text));
        ++*launch_counter();
        VK(cudaMemcpyAsync(dst, w->d_text, (size_t)nbytes, cudaMemcpyDeviceToHost, w->st));
        w->h_info[b].bytes = (unsigned long long)nbytes;
        w->h_info[b].exceptions = 0;
    } else {
        const int nblocks = (int)((cnt + kFmtThreads - 1) / kFmtThreads);
        VK(cudaMemcpyAsync(w->d_vals, w->values + first, sizeof(double) * (size_t)cnt,
                           cudaMemcpyHostToDevice, w->st));
        VK(cudaMemsetAsync(w->d_exc, 0, sizeof(unsigned int), w->st));
        // plain launches: these kernels carry no griddepcontrol.wait (launch.h)
        k_vtk_convert<<<(unsigned)nblocks, kFmtThreads, 0, w->st>>>(
            w->d_vals, cnt, w->d_packed, w->d_block_neg, w->d_exc, w->d_p10, w->d_p10e);
        k_vtk_scan<<<1, 1024, 0, w->st>>>(w->d_block_neg, nblocks, w->d_total, cnt);
        k_vtk_emit<<<(unsigned)nblocks, kFmtThreads, 0, w->st>>>(w->d_packed, cnt, w->d_block_neg,
                                                                 w->d_text);
        *launch_counter() += 3;
        VK(cudaMemcpyAsync(dst, w->d_text, (size_t)(16 * cnt), cudaMemcpyDeviceToHost, w->st));
        VK(cudaMemcpyAsync(&w->h_info[b].bytes, w->d_total, sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, w->st));
        VK(cudaMemcpyAsync(&w->h_info[b].exceptions, w->d_exc, sizeof(unsigned int),
                           cudaMemcpyDeviceToHost, w->st));
    }
    VK(cudaEventRecord(w->done[b], w->st));
    return 0;
}

}  // namespace

// writeOutputData (postprocess.h:5-47) as a stream of byte chunks: header, point lines,
// POINT_DATA header, value lines -- concatenated they are the reference's file
extern "C" int mgb_vtk_open(mgb_vtk **out, const double *values, int ni, int nj, int nk, double h,
                            int device)
{
    if (!out || !values)
        return fail("mgb_vtk_open: null argument");
    if (ni < 1 || nj < 1 || nk < 1)
        return fail("mgb_vtk_open: extents must be positive");
    int ndev = 0;
    if (mgb_device_count(&ndev))
        return 1;
    if (ndev < 1)
        return fail("no CUDA device: libmgb has no CPU fallback");
    VK(cudaSetDevice(device));
    mgb_vtk *w = new mgb_vtk;
    w->device = device;
    w->values = values;
    w->ni = ni, w->nj = nj, w->nk = nk;
    w->total = (long long)ni * nj * nk;
    w->nchunks = (w->total + kChunkPoints - 1) / kChunkPoints;
    char buf[512];
    // postprocess.h:13-19 and 37-41 (the reference's counters are int)
    snprintf(buf, sizeof buf,
             "# vtk DataFile Version 2.0\nPotential data\nASCII\nDATASET STRUCTURED_GRID\n"
             "DIMENSIONS %d %d %d\nPOINTS %d float\n", ni, nj, nk, (int)w->total);
    w->head0 = buf;
    snprintf(buf, sizeof buf, "\nPOINT_DATA %d\nSCALARS data float 1\nLOOKUP_TABLE default\n",
             (int)w->total);
    w->head1 = buf;
    // axis tables: "%10.8e" of h*i, formatted by the C library (postprocess.h:24,27,30: h * i)
    std::vector<char> tab((size_t)(ni + nj + nk) * 16, 0);
    {
        const int n[3] = {ni, nj, nk};
        size_t row = 0;
        for (int a = 0; a < 3; a++)
            for (int i = 0; i < n[a]; i++, row++) {
                char t[40];
                const int len = snprintf(t, sizeof t, "%10.8e", h * i);
                if (len != 14) {
                    delete w;
                    return fail("mgb_vtk_open: coordinate %g does not print as 14 characters", h * i);
                }
                memcpy(&tab[row * 16], t, 14);
            }
    }
    auto bail = [&](const char *what) {
        vtk_free(w);
        return fail("mgb_vtk_open: %s: %s", what, cudaGetErrorString(cudaGetLastError()));
    };
    if (cudaStreamCreate(&w->st) != cudaSuccess) return bail("stream");
    for (int b = 0; b < 2; b++) {
        if (cudaHostAlloc((void **)&w->h_buf[b], kBufBytes, cudaHostAllocDefault) != cudaSuccess)
            return bail("pinned buffer");
        if (cudaEventCreateWithFlags(&w->done[b], cudaEventDisableTiming) != cudaSuccess)
            return bail("event");
    }
    if (cudaHostAlloc((void **)&w->h_info, 2 * sizeof(mgb_vtk::Info), cudaHostAllocDefault) != cudaSuccess)
        return bail("pinned info");
    const size_t nb = (size_t)((kChunkPoints + kFmtThreads - 1) / kFmtThreads);
    if (cudaMalloc(&w->d_text, 45 * (size_t)kChunkPoints + 64) != cudaSuccess ||
        cudaMalloc(&w->d_vals, sizeof(double) * kChunkPoints) != cudaSuccess ||
        cudaMalloc(&w->d_packed, sizeof(unsigned long long) * kChunkPoints) != cudaSuccess ||
        cudaMalloc(&w->d_block_neg, sizeof(unsigned int) * nb) != cudaSuccess ||
        cudaMalloc(&w->d_exc, sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&w->d_total, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&w->d_tab, tab.size()) != cudaSuccess ||
        cudaMalloc(&w->d_p10, sizeof(kPow10Host)) != cudaSuccess ||
        cudaMalloc(&w->d_p10e, sizeof(kPow10ExpHost)) != cudaSuccess)
        return bail("device buffers");
    if (cudaMemcpy(w->d_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(w->d_p10, kPow10Host, sizeof(kPow10Host), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(w->d_p10e, kPow10ExpHost, sizeof(kPow10ExpHost), cudaMemcpyHostToDevice) != cudaSuccess)
        return bail("tables");
    if (vtk_issue(w, 0)) {
        vtk_free(w);
        return 1;
    }
    w->next_issue = 1;
    *out = w;
    return 0;
}

// next chunk of the file (valid until the following call); *n = 0: the file is complete
extern "C" int mgb_vtk_next(mgb_vtk *w, const char **bytes, long long *n)
{
    if (!w || !bytes || !n)
        return fail("mgb_vtk_next: null argument");
    VK(cudaSetDevice(w->device));
    const long long last = 2 * w->nchunks;
    if (w->next_return >= last) {
        *bytes = nullptr;
        *n = 0;
        return 0;
    }
    const long long pos = w->next_return++;
    const int b = (int)(pos & 1);
    // the other buffer is free again (the caller is done with what the previous call gave)
    if (w->next_issue < last && w->next_issue == pos + 1) {
        if (vtk_issue(w, w->next_issue))
            return 1;
        w->next_issue++;
    }
    VK(cudaEventSynchronize(w->done[b]));
    char *text = w->h_buf[b] + kHeadroom;
    long long len = (long long)w->h_info[b].bytes;
    if (pos >= w->nchunks && w->h_info[b].exceptions) {
        // the C library decides (ties through inexact powers of ten, non-finite values,
        // three-digit exponents): the whole chunk again, with the reference's own format
        const long long c = pos - w->nchunks, first = c * kChunkPoints;
        const long long cnt = std::min(kChunkPoints, w->total - first);
        // the device may still be filling the OTHER buffer; this one is ours
        len = 0;
        std::vector<char> tmp;
        tmp.resize((size_t)cnt * 32);
        for (long long i = 0; i < cnt; i++)
            len += snprintf(tmp.data() + len, 32, "%10.8e\n", w->values[first + i]);
        if ((size_t)len + kHeadroom > kBufBytes)
            return fail("mgb_vtk_next: formatted chunk does not fit");
        memcpy(text, tmp.data(), (size_t)len);
        w->host_chunks++;
    }
    const std::string *head = pos == 0 ? &w->head0 : (pos == w->nchunks ? &w->head1 : nullptr);
    if (head) {
        text -= head->size();
        memcpy(text, head->data(), head->size());
        len += (long long)head->size();
    }
    *bytes = text;
    *n = len;
    return 0;
}

// chunks that had to be formatted by the host's snprintf (exceptions), for reporting
extern "C" int mgb_vtk_host_chunks(const mgb_vtk *w, long long *chunks)
{
    if (!w || !chunks)
        return fail("mgb_vtk_host_chunks: null argument");
    *chunks = w->host_chunks;
    return 0;
}

extern "C" int mgb_vtk_close(mgb_vtk *w)
{
    vtk_free(w);
    return 0;
}
