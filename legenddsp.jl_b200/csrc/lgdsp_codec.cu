// decode_data on the GPU: the two LEGEND waveform codecs (LegendDataTypes.jl `decode_data`, call sites
// /root/reference/src/dsp_icpc.jl:313-314, src/dsp_puls.jl:103, src/dsp_sipm.jl:241), so that the host link carries the
// ENCODED bytes (a third of the raw samples) and the decoded waveforms only ever exist in HBM.
//
//   RadwareSigcompress   radware-sigcompress v1.0 (D. Radford): sections of <= 128 16-bit samples, each stored as
//                        (value - min) or (difference - min) in the fewest bits that hold the section's range, MSB first
//                        in 16-bit words; the LEGEND byte stream holds the words in big-endian order, samples are shifted by
//                        `shift` (-32768 for UInt16) before encoding.
//   ULEB128ZigZagDiff    first differences, zig-zag, unsigned LEB128 (the 32-bit presummed waveforms).
//
// Decoding a waveform is sequential in its headers only: one thread walks the section headers (<= 171 for 8192 samples),
// then every section is unpacked in parallel (fixed bit offsets; difference sections end in a warp scan).  The varint
// stream is cut at its terminator bytes with a block scan, values are decoded in parallel and prefix-summed.
// The host-side encoders are for tests, benchmarks and round trips; they restate the published algorithms independently
// of oracle/lgdsp_codec_oracle.c.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include "lgdsp_device.cuh"
#include "lgdsp_kernels.h"

namespace lgdsp {

// ---------------------------------------------------------------------------------------------------
// host-side encoders
// ---------------------------------------------------------------------------------------------------
namespace {
const int kMask[17] = {0, 1, 3, 7, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535};

struct WordWriter {
    uint8_t* out;
    long long cap_words, n = 0;
    bool ok = true;
    void put(unsigned v)
    {
        if (n >= cap_words) { ok = false; return; }
        out[2 * n] = (uint8_t)(v >> 8);   // big-endian 16-bit words
        out[2 * n + 1] = (uint8_t)v;
        ++n;
    }
};

// one signal; returns bytes written, -1: buffer too small, -2: a shifted sample leaves the int16 range
long long radware_encode_one(const uint16_t* x, int n, int shift, uint8_t* out, long long cap)
{
    std::vector<int> s(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) {
        s[i] = (int)x[i] + shift;
        if (s[i] < -32768 || s[i] > 32767) return -2;
    }
    WordWriter w{out, cap / 2};
    w.put((unsigned)n & 0xffffu);
    int j = 0;
    while (j < n) {
        // ranges of the values and of the first differences over the first <= 48 samples decide the method ...
        int vmax = s[j], vmin = s[j], dmax = -16000, dmin = 16000, len = 1;
        int i = j + 1;
        for (; i < n && i < j + 48; ++i, ++len) {
            vmax = s[i] > vmax ? s[i] : vmax;
            vmin = s[i] < vmin ? s[i] : vmin;
            const int ds = s[i] - s[i - 1];
            dmax = ds > dmax ? ds : dmax;
            dmin = ds < dmin ? ds : dmin;
        }
        const bool use_diff = !(vmax - vmin <= dmax - dmin);
        int nb = 2;
        // ... and the section grows (<= 128 samples) while the chosen quantity still fits its bit width
        if (!use_diff) {
            while (vmax - vmin > kMask[nb]) ++nb;
            for (; i < n && i < j + 128; ++i, ++len) {
                const int hi = s[i] > vmax ? s[i] : vmax, lo = s[i] < vmin ? s[i] : vmin;
                if (hi - lo > kMask[nb]) break;
                vmax = hi; vmin = lo;
            }
        } else {
            while (dmax - dmin > kMask[nb]) ++nb;
            for (; i < n && i < j + 128; ++i, ++len) {
                const int ds = s[i] - s[i - 1];
                const int hi = ds > dmax ? ds : dmax, lo = ds < dmin ? ds : dmin;
                if (hi - lo > kMask[nb]) break;
                dmax = hi; dmin = lo;
            }
        }
        w.put((unsigned)len);
        if (!use_diff) {
            w.put((unsigned)nb);
            w.put((unsigned)vmin & 0xffffu);
        } else {
            w.put((unsigned)(nb + 32));
            w.put((unsigned)s[j] & 0xffffu);
            w.put((unsigned)dmin & 0xffffu);
        }
        // bit packing, MSB first
        unsigned long long acc = 0;
        int nbits = 0;
        const int nvals = use_diff ? len - 1 : len;
        for (int k = 0; k < nvals; ++k) {
            const unsigned v = use_diff ? (unsigned)(s[j + 1 + k] - s[j + k] - dmin) : (unsigned)(s[j + k] - vmin);
            acc = (acc << nb) | v;
            nbits += nb;
            while (nbits >= 16) {
                w.put((unsigned)(acc >> (nbits - 16)) & 0xffffu);
                nbits -= 16;
            }
        }
        if (nbits > 0) w.put((unsigned)(acc << (16 - nbits)) & 0xffffu);
        j += len;
    }
    if (w.n % 2) w.put(0);   // 4-byte granularity
    return w.ok ? 2 * w.n : -1;
}

long long uleb_encode_one(const void* x, int sample_bytes, int n, uint8_t* out, long long cap)
{
    long long pos = 0, last = 0;
    for (int i = 0; i < n; ++i) {
        const long long v = sample_bytes == 4 ? (long long)static_cast<const uint32_t*>(x)[i] : (long long)static_cast<const uint16_t*>(x)[i];
        const long long d = v - last;
        last = v;
        unsigned long long z = ((unsigned long long)d << 1) ^ (unsigned long long)(d >> 63);
        do {
            if (pos >= cap) return -1;
            uint8_t b = (uint8_t)(z & 0x7f);
            z >>= 7;
            out[pos++] = z ? (uint8_t)(b | 0x80) : b;
        } while (z);
    }
    return pos;
}
}  // namespace

long long codec_max_encoded_bytes(int codec, int n_samples, int sample_bytes)
{
    if (codec == LGDSP_CODEC_RADWARE) return 2 + (long long)((n_samples + 47) / 48) * 8 + 2LL * n_samples + 6;
    return (long long)n_samples * (sample_bytes == 4 ? 5 : 3);
}

// encodes n_events waveforms into one contiguous buffer; offsets[e] .. offsets[e+1] delimit event e.  Two passes: every
// worker encodes its events into a private worst-case buffer, then the streams are packed.
int codec_encode_host(int codec, const void* wf, int sample_bytes, long long n_events, int n_samples, long long ld, int shift,
                      uint8_t* enc, long long cap, long long* offsets)
{
    const long long worst = codec_max_encoded_bytes(codec, n_samples, sample_bytes);
    unsigned nthr = std::thread::hardware_concurrency();
    if (nthr < 1) nthr = 1;
    if ((long long)nthr > n_events) nthr = (unsigned)(n_events > 0 ? n_events : 1);
    std::vector<std::vector<uint8_t>> bufs(nthr);
    std::vector<long long> sizes((size_t)n_events, 0);
    std::vector<int> rcs(nthr, 0);
    auto work = [&](unsigned t) {
        const long long e0 = n_events * t / nthr, e1 = n_events * (t + 1) / nthr;
        std::vector<uint8_t>& b = bufs[t];
        b.resize((size_t)((e1 - e0) * worst + 16));
        long long pos = 0;
        for (long long e = e0; e < e1; ++e) {
            const uint8_t* src = static_cast<const uint8_t*>(wf) + (size_t)e * (size_t)ld * (size_t)sample_bytes;
            const long long nbytes = codec == LGDSP_CODEC_RADWARE
                                         ? radware_encode_one(reinterpret_cast<const uint16_t*>(src), n_samples, shift, b.data() + pos, worst)
                                         : uleb_encode_one(src, sample_bytes, n_samples, b.data() + pos, worst);
            if (nbytes < 0) { rcs[t] = (int)nbytes; return; }
            sizes[(size_t)e] = nbytes;
            pos += nbytes;
        }
        b.resize((size_t)pos);
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nthr; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    for (unsigned t = 0; t < nthr; ++t)
        if (rcs[t]) return rcs[t];
    offsets[0] = 0;
    for (long long e = 0; e < n_events; ++e) offsets[e + 1] = offsets[e] + sizes[(size_t)e];
    if (offsets[n_events] > cap) return -1;
    for (unsigned t = 0; t < nthr; ++t) {
        const long long e0 = n_events * t / nthr;
        if (!bufs[t].empty()) memcpy(enc + offsets[e0], bufs[t].data(), bufs[t].size());
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// device decoders
// ---------------------------------------------------------------------------------------------------
constexpr int DEC_NT = 128;
// bytes of one stream -> shared memory; returns the index of stream byte 0 inside `sm` (the copy is word aligned: whole
// 32-bit words where they lie inside the stream, single bytes at its two ends -- nothing outside the stream is read)
__device__ __forceinline__ int load_stream(const uint8_t* __restrict__ enc, long long off, long long nbytes, unsigned char* sm, int tid)
{
    const uint8_t* p = enc + off;
    const int skew = (int)(reinterpret_cast<uintptr_t>(p) & 3);
    const int head = skew ? min(4 - skew, (int)nbytes) : 0;          // bytes in front of the first aligned word
    const int nwords = (int)((nbytes - head) >> 2);
    const int tail = (int)(nbytes - head - 4LL * nwords);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p + head);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sm + skew + head);   // (skew + head is 0 or 4)
    for (int i = tid; i < nwords; i += DEC_NT) dst[i] = __ldg(src + i);
    if (tid < head) sm[skew + tid] = __ldg(p + tid);
    if (tid < tail) sm[skew + head + 4 * nwords + tid] = __ldg(p + head + 4 * nwords + tid);
    return skew;
}

// One stream -> the warp's shared-memory buffer with stream byte 0 on a 4-byte boundary, 16 zero bytes behind it.  Whole
// aligned 32-bit words where they lie inside the stream (funnel shift when the stream starts between word boundaries),
// single bytes at the ends: nothing outside the stream is read.
__device__ __forceinline__ void load_stream_warp(const uint8_t* __restrict__ p, int nbytes, uint32_t* dst, int lane)
{
    const int skew = (int)(reinterpret_cast<uintptr_t>(p) & 3);
    const int nfull = nbytes >> 2, tail = nbytes & 3;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(p - skew);   // a[i], a[i+1] cover destination word i
    if (skew == 0) {
        int i = lane;
        for (; i + 96 < nfull; i += 128) {   // four loads in flight
            const uint32_t v0 = __ldg(a + i), v1 = __ldg(a + i + 32), v2 = __ldg(a + i + 64), v3 = __ldg(a + i + 96);
            dst[i] = v0; dst[i + 32] = v1; dst[i + 64] = v2; dst[i + 96] = v3;
        }
        for (; i < nfull; i += 32) dst[i] = __ldg(a + i);
    } else {
        for (int i = lane; i < nfull; i += 32) {
            uint32_t v;
            if (i >= 1 && 4 * (i + 2) - skew <= nbytes) v = __funnelshift_r(__ldg(a + i), __ldg(a + i + 1), 8 * skew);
            else {
                const uint8_t* q = p + 4 * i;
                v = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16) | ((uint32_t)__ldg(q + 3) << 24);
            }
            dst[i] = v;
        }
    }
    if (lane < 4) {
        uint32_t v = 0;
        if (lane == 0) for (int t = 0; t < tail; ++t) v |= (uint32_t)__ldg(p + 4 * nfull + t) << (8 * t);
        dst[nfull + lane] = v;
    }
}

// out[e][i] = sample i of event e (uint16); status[e] = 0 ok, 1 malformed stream (the waveform is zero-filled); err[0] counts
// the malformed streams, err[1] keeps the smallest err_base + e among them (both optional).
// One WARP per waveform: it walks the sections in order (the header fields are read by every lane: shared-memory
// broadcasts, no divergence, no block barrier) and unpacks a section's <= 128 values four per lane.  `cap_w` = bytes of
// one warp's stream buffer (>= longest stream + 16).
constexpr int DEC_WARPS = DEC_NT / 32;
__global__ void __launch_bounds__(DEC_NT)
radware_decode_kernel(const uint8_t* __restrict__ enc, const long long* __restrict__ off, long long off_base, long long n_events,
                      int n_samples, int shift, uint16_t* __restrict__ out, long long ld, int* __restrict__ status, int* __restrict__ err,
                      int err_base, int cap_w)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* w32 = reinterpret_cast<uint32_t*>(dsm + (size_t)wid * cap_w);
    const uint16_t* h16 = reinterpret_cast<const uint16_t*>(w32);
    auto word = [&](int k) -> int { return (int)__byte_perm((unsigned)h16[k], 0u, 0x4401); };   // big-endian 16-bit word k
    auto bswap = [](uint32_t v) -> uint32_t { return __byte_perm(v, 0u, 0x0123); };
    for (long long e = (long long)blockIdx.x * DEC_WARPS + wid; e < n_events; e += (long long)gridDim.x * DEC_WARPS) {
        const long long o0 = off[e] - off_base, nbytes = off[e + 1] - off[e];
        int bad = (nbytes >= 2 && nbytes + 16 <= cap_w) ? 0 : 1;
        uint16_t* o = out + e * ld;
        __syncwarp();   // every lane is done with the previous stream
        if (!bad) {
            load_stream_warp(enc + o0, (int)nbytes, w32, lane);
            __syncwarp();
            const int nwords = (int)(nbytes >> 1);
            const int siglen = word(0);
            if (siglen != n_samples) bad = 1;
            int pos = 1, iso = 0;
            while (!bad && pos < nwords && iso < siglen) {
                if (pos + 3 > nwords) { bad = 1; break; }
                int nw = word(pos), nb = word(pos + 1), start = 0, mn, payload;
                const bool diff = nb >= 32;
                if (diff) {
                    nb -= 32;
                    if (pos + 4 > nwords) { bad = 1; break; }
                    start = (short)word(pos + 2);
                    mn = (short)word(pos + 3);
                    payload = pos + 4;
                } else {
                    mn = (short)word(pos + 2);
                    payload = pos + 3;
                }
                if (nb > 16 || nw < 1) { bad = 1; break; }
                const int nvals = diff ? nw - 1 : nw;
                const int pw = (nvals * nb + 15) >> 4;
                if (payload + pw > nwords) { bad = 1; break; }
                if (nw > siglen - iso) nw = siglen - iso;   // the original stops at the stored length
                const int nv = diff ? nw - 1 : nw;          // values to unpack
                uint16_t* os = o + iso + (diff ? 1 : 0);
                int carry = start;                           // running value in front of this round (16-bit wrap-around at the store)
                if (diff && lane == 0) o[iso] = (uint16_t)((carry - shift) & 0xffff);
                const int rsh = 32 - nb;
                for (int k0 = 0; k0 < nv; k0 += 128) {
                    // this lane's four values start at bit (k0 + 4 lane) nb of the payload: at most 16 + 15 + 64 bits = three words
                    const int kb = k0 + 4 * lane;
                    int v[4] = {0, 0, 0, 0};
                    if (kb < nv && nb > 0) {
                        const int p0 = kb * nb;
                        const int hw = payload + (p0 >> 4);
                        const uint32_t* q = w32 + (hw >> 1);
                        const uint32_t B0 = bswap(q[0]), B1 = bswap(q[1]), B2 = bswap(q[2]);
                        int t = ((hw & 1) << 4) + (p0 & 15);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int j = t >> 5;
                            const uint32_t hi = j == 0 ? B0 : (j == 1 ? B1 : B2), lo = j == 0 ? B1 : (j == 1 ? B2 : 0u);
                            v[r] = (int)(__funnelshift_l(lo, hi, t & 31) >> rsh);
                            t += nb;
                        }
                    }
                    if (!diff) {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (kb + r < nv) os[kb + r] = (uint16_t)((v[r] + mn - shift) & 0xffff);
                    } else {
                        int sum = 0;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            sum += (kb + r < nv) ? v[r] + mn : 0;
                            v[r] = sum;
                        }
                        int incl = sum;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const int t = __shfl_up_sync(FULL, incl, d);
                            if (lane >= d) incl += t;
                        }
                        const int base = carry + incl - sum;
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (kb + r < nv) os[kb + r] = (uint16_t)((base + v[r] - shift) & 0xffff);
                        carry += __shfl_sync(FULL, incl, 31);
                    }
                }
                iso += nw;
                pos = payload + pw;
            }
            if (!bad && iso != siglen) bad = 1;
        }
        if (bad) {
            __syncwarp();
            for (int i = lane; i < n_samples; i += 32) o[i] = 0;
        }
        if (lane == 0) {
            if (status) status[e] = bad;
            if (bad && err) { atomicAdd(&err[0], 1); atomicMin(&err[1], err_base + (int)e); }
        }
    }
}

// ULEB128 zig-zag differences -> uint16 / uint32 samples
#define VPAD(i) ((i) + ((i) >> 5))   // table index with one pad word per 32: the per-thread runs start in different banks
template <typename OUT>
__global__ void __launch_bounds__(DEC_NT)
uleb_decode_kernel(const uint8_t* __restrict__ enc, const long long* __restrict__ off, long long off_base, long long n_events,
                   int n_samples, OUT* __restrict__ out, long long ld, int* __restrict__ status, int* __restrict__ err, int err_base,
                   int cap_bytes)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    unsigned char* bytes = dsm;                                                             // cap_bytes + 16
    uint32_t* vals = reinterpret_cast<uint32_t*>(dsm + ((cap_bytes + 16 + 15) & ~15));      // n_samples differences
    __shared__ int s_scan[DEC_NT / 32 + 1];
    __shared__ uint32_t s_tot[DEC_NT / 32 + 1];
    __shared__ int s_err;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (long long e = blockIdx.x; e < n_events; e += gridDim.x) {
        const long long o0 = off[e] - off_base, nbytes = off[e + 1] - off[e];
        const bool fits = nbytes >= 0 && nbytes + 4 <= cap_bytes;
        int skew = 0;
        if (tid == 0) s_err = fits ? 0 : 1;
        if (fits) skew = load_stream(enc, o0, nbytes, bytes, tid);
        __syncthreads();
        const unsigned char* b = bytes + skew;
        const int nb = fits ? (int)nbytes : 0;
        // values START at byte 0 and behind every terminator byte (bit 7 clear): count the starts of this thread's byte range
        const int per = (nb + DEC_NT - 1) / DEC_NT;
        const int lo = tid * per, hi = min(lo + per, nb);
        int cnt = 0;
        for (int p = lo; p < hi; ++p) cnt += (p == 0 || b[p - 1] < 0x80) ? 1 : 0;
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_scan[wid] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
        for (int w = 0; w < DEC_NT / 32; ++w) { if (w < wid) wbase += s_scan[w]; total += s_scan[w]; }
        int idx = wbase + incl - cnt;
        if (total != n_samples || (nb > 0 && b[nb - 1] >= 0x80)) { if (tid == 0) s_err = 1; }
        else {
            for (int p = lo; p < hi; ++p) {
                if (!(p == 0 || b[p - 1] < 0x80)) continue;
                unsigned long long z = 0;
                int sh = 0, q = p;
                for (;;) {
                    const unsigned c = b[q++];
                    z |= (unsigned long long)(c & 0x7f) << sh;
                    sh += 7;
                    if (c < 0x80 || q >= nb || sh > 63) break;
                }
                vals[VPAD(idx)] = (uint32_t)((z >> 1) ^ (0ull - (z & 1ull)));   // zig-zag decode; the sums wrap in 32 bits
                ++idx;
            }
        }
        __syncthreads();
        OUT* o = out + e * ld;
        if (s_err) {
            for (int i = tid; i < n_samples; i += DEC_NT) o[i] = 0;
        } else {
            // prefix sum of the differences: every thread sums a contiguous run, block scan of the run totals; the sums go
            // back into the table and leave through a coalesced copy
            const int run = (n_samples + DEC_NT - 1) / DEC_NT;
            const int a = tid * run, z = min(a + run, n_samples);
            uint32_t sum = 0;
            for (int i = a; i < z; ++i) sum += vals[VPAD(i)];
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) s_tot[wid] = inc;
            __syncthreads();
            uint32_t wb = 0;
            for (int w = 0; w < wid; ++w) wb += s_tot[w];
            uint32_t acc = wb + inc - sum;
            for (int i = a; i < z; ++i) {
                acc += vals[VPAD(i)];
                vals[VPAD(i)] = acc;
            }
            __syncthreads();
            if (sizeof(OUT) == 2 && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
                uint32_t* o2 = reinterpret_cast<uint32_t*>(o);
                for (int i = tid; 2 * i + 1 < n_samples; i += DEC_NT)
                    o2[i] = (vals[VPAD(2 * i)] & 0xffffu) | (vals[VPAD(2 * i + 1)] << 16);
                if ((n_samples & 1) && tid == 0) o[n_samples - 1] = (OUT)vals[VPAD(n_samples - 1)];
            } else {
                for (int i = tid; i < n_samples; i += DEC_NT) o[i] = (OUT)vals[VPAD(i)];
            }
        }
        if (tid == 0) {
            if (status) status[e] = s_err;
            if (s_err && err) { atomicAdd(&err[0], 1); atomicMin(&err[1], err_base + (int)e); }
        }
        __syncthreads();
    }
}

static int dec_smem(int codec, int n_samples, int sample_bytes, long long max_stream_bytes, int* cap_bytes)
{
    const long long worst = codec_max_encoded_bytes(codec, n_samples, sample_bytes);
    if (codec == LGDSP_CODEC_RADWARE) {
        // per-warp stream buffers sized for the longest stream of the batch when the caller knows it (host offsets)
        const long long m = (max_stream_bytes > 0 && max_stream_bytes < worst) ? max_stream_bytes : worst;
        const int cap = (int)((m + 16 + 15) & ~15LL);
        *cap_bytes = cap;
        return cap * DEC_WARPS;
    }
    const int cap = (int)worst + 8;
    *cap_bytes = cap;
    return ((cap + 16 + 15) & ~15) + (n_samples + (n_samples >> 5) + 1) * 4;
}

// d_off[e] - off_base = first byte of event e inside d_enc (off_base = d_off[0] when only a slice of the bytes was uploaded);
// max_stream_bytes = longest single stream of the batch if known (0: the codec's worst case)
cudaError_t codec_decode_launch(int codec, const uint8_t* d_enc, const long long* d_off, long long off_base, long long n_events,
                                int n_samples, int shift, void* d_out, int sample_bytes, long long ld, int* d_status, int* d_err,
                                int err_base, long long max_stream_bytes, int sm_count, cudaStream_t stream)
{
    if (n_events <= 0) return cudaSuccess;
    int cap = 0;
    const int smem = dec_smem(codec, n_samples, sample_bytes, max_stream_bytes, &cap);
    const long long maxg = (long long)sm_count * 16;
    cudaError_t err;
    if (codec == LGDSP_CODEC_RADWARE) {
        const long long want = (n_events + DEC_WARPS - 1) / DEC_WARPS;
        const int grid = (int)(want < maxg ? want : maxg);
        if ((err = cudaFuncSetAttribute(radware_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return err;
        radware_decode_kernel<<<grid, DEC_NT, smem, stream>>>(d_enc, d_off, off_base, n_events, n_samples, shift, static_cast<uint16_t*>(d_out), ld,
                                                               d_status, d_err, err_base, cap);
        return cudaGetLastError();
    }
    const int grid = (int)(n_events < maxg ? n_events : maxg);
    if (sample_bytes == 4) {
        if ((err = cudaFuncSetAttribute(uleb_decode_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return err;
        uleb_decode_kernel<uint32_t><<<grid, DEC_NT, smem, stream>>>(d_enc, d_off, off_base, n_events, n_samples, static_cast<uint32_t*>(d_out), ld,
                                                                     d_status, d_err, err_base, cap);
    } else {
        if ((err = cudaFuncSetAttribute(uleb_decode_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return err;
        uleb_decode_kernel<uint16_t><<<grid, DEC_NT, smem, stream>>>(d_enc, d_off, off_base, n_events, n_samples, static_cast<uint16_t*>(d_out), ld,
                                                                     d_status, d_err, err_base, cap);
    }
    return cudaGetLastError();
}

}  // namespace lgdsp
