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
constexpr int RW_MAXSEC = 512;   // sections per waveform the table holds (the encoder makes >= 48-sample sections: <= 171)

struct RwSec {
    int iso, nw, nb, diff, mn, start, payload;   // payload: first payload word of the section
};

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

// out[e][i] = sample i of event e (uint16); status[e] = 0 ok, 1 malformed stream (the waveform is zero-filled); err[0] counts
// the malformed streams, err[1] keeps the smallest err_base + e among them (both optional)
__global__ void __launch_bounds__(DEC_NT)
radware_decode_kernel(const uint8_t* __restrict__ enc, const long long* __restrict__ off, long long off_base, long long n_events,
                      int n_samples, int shift, uint16_t* __restrict__ out, long long ld, int* __restrict__ status, int* __restrict__ err,
                      int err_base, int cap_bytes)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    unsigned char* bytes = dsm;                                                   // cap_bytes + 8
    RwSec* sec = reinterpret_cast<RwSec*>(dsm + ((cap_bytes + 8 + 15) & ~15));    // RW_MAXSEC
    __shared__ int s_nsec, s_err;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (long long e = blockIdx.x; e < n_events; e += gridDim.x) {
        const long long o0 = off[e] - off_base, nbytes = off[e + 1] - off[e];
        const bool fits = nbytes >= 2 && nbytes + 4 <= cap_bytes;
        int skew = 0;
        if (fits) skew = load_stream(enc, o0, nbytes, bytes, tid);
        __syncthreads();
        const unsigned char* b = bytes + skew;
        auto word = [&](int k) -> unsigned { return ((unsigned)b[2 * k] << 8) | b[2 * k + 1]; };
        if (tid == 0) {
            // the section headers are the only sequential part
            int err = fits ? 0 : 1, ns = 0;
            if (!err) {
                const int nwords = (int)(nbytes >> 1);
                const int siglen = (int)word(0);
                if (siglen != n_samples) err = 1;
                int pos = 1, iso = 0;
                while (!err && pos < nwords && iso < siglen) {
                    if (ns >= RW_MAXSEC || pos + 3 > nwords) { err = 1; break; }
                    RwSec s;
                    s.iso = iso;
                    s.nw = (int)word(pos);
                    int nb = (int)word(pos + 1);
                    s.diff = nb >= 32;
                    if (s.diff) {
                        nb -= 32;
                        if (pos + 4 > nwords) { err = 1; break; }
                        s.start = (short)word(pos + 2);
                        s.mn = (short)word(pos + 3);
                        s.payload = pos + 4;
                    } else {
                        s.start = 0;
                        s.mn = (short)word(pos + 2);
                        s.payload = pos + 3;
                    }
                    s.nb = nb;
                    if (nb > 16 || s.nw < 1) { err = 1; break; }
                    const int nvals = s.diff ? s.nw - 1 : s.nw;
                    const int pw = (nvals * nb + 15) >> 4;
                    if (s.payload + pw > nwords) { err = 1; break; }
                    if (s.nw > siglen - iso) s.nw = siglen - iso;   // the original stops at the stored length
                    sec[ns++] = s;
                    iso += s.nw;
                    pos = s.payload + pw;
                }
                if (!err && iso != siglen) err = 1;
            }
            s_nsec = ns;
            s_err = err;
        }
        __syncthreads();
        uint16_t* o = out + e * ld;
        if (s_err) {
            for (int i = tid; i < n_samples; i += DEC_NT) o[i] = 0;
        } else {
            // one warp per section; a lane unpacks four consecutive values per round of 128
            for (int si = wid; si < s_nsec; si += DEC_NT / 32) {
                const RwSec s = sec[si];
                const unsigned mask = s.nb ? (0xffffffffu >> (32 - s.nb)) : 0u;
                auto val = [&](int k) -> int {   // payload value k
                    const int p = k * s.nb, wq = s.payload + (p >> 4), sh = p & 15;
                    const unsigned x = (word(wq) << 16) | word(wq + 1);   // (one word of slack behind the stream is zero padded)
                    return s.nb ? (int)((x >> (32 - sh - s.nb)) & mask) + s.mn : s.mn;
                };
                if (!s.diff) {
                    for (int k = lane; k < s.nw; k += 32) o[s.iso + k] = (uint16_t)((val(k) - shift) & 0xffff);
                } else {
                    int carry = s.start;   // running value in front of this round (16-bit wrap-around is applied at the store)
                    if (lane == 0) o[s.iso] = (uint16_t)((carry - shift) & 0xffff);
                    for (int k0 = 0; k0 < s.nw - 1; k0 += 128) {
                        int v[4], sum = 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int k = k0 + 4 * lane + q;
                            v[q] = k < s.nw - 1 ? val(k) : 0;
                            sum += v[q];
                            v[q] = sum;
                        }
                        int incl = sum;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const int t = __shfl_up_sync(FULL, incl, d);
                            if (lane >= d) incl += t;
                        }
                        const int base = carry + incl - sum;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int k = k0 + 4 * lane + q;
                            if (k < s.nw - 1) o[s.iso + 1 + k] = (uint16_t)((base + v[q] - shift) & 0xffff);
                        }
                        carry += __shfl_sync(FULL, incl, 31);
                    }
                }
            }
        }
        if (tid == 0) {
            if (status) status[e] = s_err;
            if (s_err && err) { atomicAdd(&err[0], 1); atomicMin(&err[1], err_base + (int)e); }
        }
        __syncthreads();   // the byte buffer and the section table are reused by the next event
    }
}

// ULEB128 zig-zag differences -> uint16 / uint32 samples
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
                vals[idx++] = (uint32_t)((z >> 1) ^ (0ull - (z & 1ull)));   // zig-zag decode; the sums wrap in 32 bits
            }
        }
        __syncthreads();
        OUT* o = out + e * ld;
        if (s_err) {
            for (int i = tid; i < n_samples; i += DEC_NT) o[i] = 0;
        } else {
            // prefix sum of the differences: every thread sums a contiguous run, block scan of the run totals
            const int run = (n_samples + DEC_NT - 1) / DEC_NT;
            const int a = tid * run, z = min(a + run, n_samples);
            uint32_t sum = 0;
            for (int i = a; i < z; ++i) sum += vals[i];
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
                acc += vals[i];
                o[i] = (OUT)acc;
            }
        }
        if (tid == 0) {
            if (status) status[e] = s_err;
            if (s_err && err) { atomicAdd(&err[0], 1); atomicMin(&err[1], err_base + (int)e); }
        }
        __syncthreads();
    }
}

static int dec_smem(int codec, int n_samples, int sample_bytes, int* cap_bytes)
{
    const int cap = (int)codec_max_encoded_bytes(codec, n_samples, sample_bytes) + 8;
    *cap_bytes = cap;
    if (codec == LGDSP_CODEC_RADWARE) return ((cap + 8 + 15) & ~15) + RW_MAXSEC * (int)sizeof(RwSec);
    return ((cap + 16 + 15) & ~15) + n_samples * 4;
}

// d_off[e] - off_base = first byte of event e inside d_enc (off_base = d_off[0] when only a slice of the bytes was uploaded)
cudaError_t codec_decode_launch(int codec, const uint8_t* d_enc, const long long* d_off, long long off_base, long long n_events,
                                int n_samples, int shift, void* d_out, int sample_bytes, long long ld, int* d_status, int* d_err,
                                int err_base, int sm_count, cudaStream_t stream)
{
    if (n_events <= 0) return cudaSuccess;
    int cap = 0;
    const int smem = dec_smem(codec, n_samples, sample_bytes, &cap);
    const long long maxg = (long long)sm_count * 16;
    const int grid = (int)(n_events < maxg ? n_events : maxg);
    cudaError_t err;
    if (codec == LGDSP_CODEC_RADWARE) {
        if ((err = cudaFuncSetAttribute(radware_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return err;
        radware_decode_kernel<<<grid, DEC_NT, smem, stream>>>(d_enc, d_off, off_base, n_events, n_samples, shift, static_cast<uint16_t*>(d_out), ld,
                                                               d_status, d_err, err_base, cap);
    } else if (sample_bytes == 4) {
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
