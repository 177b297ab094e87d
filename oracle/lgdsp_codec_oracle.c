/*
 * lgdsp_codec_oracle.c -- CPU restatement of the two waveform codecs behind `decode_data`.  TEST INFRASTRUCTURE ONLY
 * (same rules as lgdsp_oracle.c: only tests/, smoke() and bench.py's CPU legs may load this library).
 *
 * Call sites in the reference: /root/reference/src/dsp_icpc.jl:313-314 (`decode_data(data.waveform_presummed)`,
 * `decode_data(data.waveform_windowed)`), src/dsp_puls.jl:103, src/dsp_sipm.jl:241.  `decode_data` and the codecs live in
 * LegendDataTypes.jl (Project.toml:13, compat "0.1.13", no Manifest: not vendored under /root/reference), so this file
 * restates the PUBLISHED algorithms:
 *
 *   RadwareSigcompress    radware-sigcompress v1.0 (D. Radford, ORNL, `compress_signal` / `decompress_signal` of
 *                         sigcompress.c); LegendDataTypes' `RadwareSigcompress(shift)` and legend-pydataobj's
 *                         `lgdo.compression.radware` are ports of it.  16-bit samples; a signal is cut into sections of
 *                         <= 128 samples, each stored either as (value - min) or as (difference - min) in the fewest
 *                         bits that hold the section's range, MSB first in 16-bit words.  The byte stream of the LEGEND
 *                         ports holds those 16-bit words in big-endian ("network") order and the samples are shifted by
 *                         `shift` before encoding (-32768 for UInt16 input) [recalled from the ports; the word order is
 *                         a parameter below so that a file-based check can flip it].
 *   ULEB128ZigZagDiff     legend-pydataobj `lgdo.compression.varlen` / LegendDataTypes' VarlenDiffArrayCodec: first
 *                         differences (x[-1] = 0), zig-zag mapped to unsigned, unsigned LEB128 (7 bits per byte, low
 *                         group first, bit 7 = continuation).  Used for the 32-bit presummed waveforms.
 *
 * PARITY STATUS: "parity unpinned" against LegendDataTypes itself (no Julia, no LH5 fixture in the reference tree).  Pinned
 * by construction properties instead: hand-worked vectors (tests/golden/make_codec_kat.py), encode -> decode round trips
 * on every population, and agreement with the product's independent host encoders.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static const int RW_MASK[17] = {0, 1, 3, 7, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535};

static inline void put16(uint8_t* out, int64_t word, uint32_t v, int big_endian)
{
    if (big_endian) { out[2 * word] = (uint8_t)(v >> 8); out[2 * word + 1] = (uint8_t)v; }
    else { out[2 * word] = (uint8_t)v; out[2 * word + 1] = (uint8_t)(v >> 8); }
}
static inline uint32_t get16(const uint8_t* in, int64_t word, int big_endian)
{
    return big_endian ? ((uint32_t)in[2 * word] << 8) | in[2 * word + 1] : ((uint32_t)in[2 * word + 1] << 8) | in[2 * word];
}

/* compress_signal (sigcompress.c): returns the number of BYTES written, -1 if `cap` bytes do not suffice, -2 if a shifted
 * sample leaves the int16 range.  Worst case: 2 + ceil(n/48) * 8 + 2 n + 4 bytes. */
ORC_API int64_t orc_radware_encode(const uint16_t* x, int n, int shift, int big_endian, uint8_t* out, int64_t cap)
{
    int16_t* s = (int16_t*)malloc(sizeof(int16_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        const int v = (int)x[i] + shift;
        if (v < -32768 || v > 32767) { free(s); return -2; }
        s[i] = (int16_t)v;
    }
    const int64_t capw = cap / 2;
    int64_t iso = 0;
    int j = 0;
#define NEED(w) do { if ((w) > capw) { free(s); return -1; } } while (0)
    NEED(1);
    put16(out, iso++, (uint32_t)n & 0xffffu, big_endian);   /* signal length */
    while (j < n) {
        /* find the method and length of the next section */
        int max1 = s[j], min1 = s[j], max2 = -16000, min2 = 16000, nb1 = 2, nb2 = 2, nw = 1, i;
        for (i = j + 1; i < n && i < j + 48; ++i) {
            if (max1 < s[i]) max1 = s[i];
            if (min1 > s[i]) min1 = s[i];
            const int ds = s[i] - s[i - 1];
            if (max2 < ds) max2 = ds;
            if (min2 > ds) min2 = ds;
            ++nw;
        }
        if (max1 - min1 <= max2 - min2) {   /* absolute values */
            nb2 = 99;
            while (max1 - min1 > RW_MASK[nb1]) ++nb1;   /* (signed compare, as in the original) */
            for (; i < n && i < j + 128; ++i) {
                if (max1 < s[i]) max1 = s[i];
                int dd1 = max1 - min1;
                if (min1 > s[i]) dd1 = max1 - s[i];
                if (dd1 > RW_MASK[nb1]) break;
                if (min1 > s[i]) min1 = s[i];
                ++nw;
            }
        } else {                            /* differences */
            nb1 = 99;
            while (max2 - min2 > RW_MASK[nb2]) ++nb2;   /* a one-sample section has max2 - min2 = -32000: nb2 stays 2 */
            for (; i < n && i < j + 128; ++i) {
                const int ds = s[i] - s[i - 1];
                if (max2 < ds) max2 = ds;
                int dd2 = max2 - min2;
                if (min2 > ds) dd2 = max2 - ds;
                if (dd2 > RW_MASK[nb2]) break;
                if (min2 > ds) min2 = ds;
                ++nw;
            }
        }
        /* the section: header words, then the values MSB first in 16-bit words */
        const int diff = nb1 > nb2;
        const int nb = diff ? nb2 : nb1, nvals = diff ? nw - 1 : nw;
        const int64_t words = ((int64_t)nvals * nb + 15) / 16;
        NEED(iso + (diff ? 4 : 3) + words);
        put16(out, iso++, (uint32_t)nw, big_endian);
        if (!diff) {
            put16(out, iso++, (uint32_t)nb1, big_endian);
            put16(out, iso++, (uint32_t)min1 & 0xffffu, big_endian);
        } else {
            put16(out, iso++, (uint32_t)(nb2 + 32), big_endian);
            put16(out, iso++, (uint32_t)s[j] & 0xffffu, big_endian);
            put16(out, iso++, (uint32_t)min2 & 0xffffu, big_endian);
        }
        uint32_t acc = 0;   /* bits waiting in the current 16-bit word, left aligned in 32 bits */
        int bp = 0;
        int64_t w = iso;
        for (int k = 0; k < nvals; ++k) {
            const uint32_t v = diff ? (uint32_t)(s[j + 1 + k] - s[j + k] - min2) : (uint32_t)(s[j + k] - min1);
            acc |= v << (32 - bp - nb);
            bp += nb;
            if (bp > 15) {
                put16(out, w++, acc >> 16, big_endian);
                acc <<= 16;
                bp -= 16;
            }
        }
        if (bp > 0) put16(out, w++, acc >> 16, big_endian);
        iso += words;
        j += nw;
    }
    if (iso % 2) { NEED(iso + 1); put16(out, iso++, 0, big_endian); }   /* 4-byte padding */
#undef NEED
    free(s);
    return 2 * iso;
}

/* decompress_signal (sigcompress.c): returns the number of samples written (the stored signal length), -1 on a malformed
 * stream (sections beyond the buffer, bit width > 16, more samples than `cap`) */
ORC_API int orc_radware_decode(const uint8_t* in, int64_t nbytes, int shift, int big_endian, uint16_t* x, int cap)
{
    const int64_t nwords = nbytes / 2;
    if (nwords < 1) return -1;
    int64_t isi = 0;
    const int siglen = (int)get16(in, isi++, big_endian);
    if (siglen > cap) return -1;
    int iso = 0;
    while (isi < nwords && iso < siglen) {
        if (isi + 3 > nwords) return -1;
        const int nw = (int)get16(in, isi++, big_endian);
        int nb = (int)get16(in, isi++, big_endian);
        int prev = 0, first = 0, nvals = nw;
        const int diff = nb >= 32;
        if (diff) {
            nb -= 32;
            if (isi + 2 > nwords) return -1;
            prev = (int16_t)get16(in, isi++, big_endian);   /* starting value */
            x[iso++] = (uint16_t)((prev - shift) & 0xffff);
            first = 1;
            nvals = nw - 1;
        }
        if (nb > 16 || nw < 1) return -1;
        const int mn = (int16_t)get16(in, isi++, big_endian);
        const int64_t words = ((int64_t)nvals * nb + 15) / 16;
        if (isi + words > nwords) return -1;
        (void)first;
        for (int k = 0; k < nvals && iso < siglen; ++k) {
            const int64_t p = (int64_t)k * nb;
            const int64_t w = isi + (p >> 4);
            const int o = (int)(p & 15);
            const uint32_t hi = get16(in, w, big_endian), lo = (w + 1 < nwords) ? get16(in, w + 1, big_endian) : 0u;
            const uint32_t v = nb ? ((((hi << 16) | lo) >> (32 - o - nb)) & (uint32_t)RW_MASK[nb]) : 0u;
            int val = (int)v + mn;
            if (diff) { val = (int16_t)(val + prev); }   /* the original accumulates in 16-bit shorts */
            prev = val;
            x[iso++] = (uint16_t)(((int16_t)val - shift) & 0xffff);
        }
        isi += words;
    }
    return iso == siglen ? siglen : -1;
}

/* ULEB128 zig-zag difference codec (lgdo.compression.varlen): n unsigned samples of `sample_bytes` (2 or 4) bytes */
ORC_API int64_t orc_uleb128zzd_encode(const void* x, int sample_bytes, int n, uint8_t* out, int64_t cap)
{
    int64_t pos = 0, last = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t v = sample_bytes == 4 ? (int64_t)((const uint32_t*)x)[i] : (int64_t)((const uint16_t*)x)[i];
        const int64_t d = v - last;
        last = v;
        uint64_t z = ((uint64_t)d << 1) ^ (uint64_t)(d >> 63);
        do {
            if (pos >= cap) return -1;
            uint8_t b = (uint8_t)(z & 0x7f);
            z >>= 7;
            if (z) b |= 0x80;
            out[pos++] = b;
        } while (z);
    }
    return pos;
}

ORC_API int orc_uleb128zzd_decode(const uint8_t* in, int64_t nbytes, int sample_bytes, void* x, int cap)
{
    int64_t pos = 0, last = 0;
    int n = 0;
    while (pos < nbytes) {
        uint64_t z = 0;
        int sh = 0;
        for (;;) {
            if (pos >= nbytes || sh > 63) return -1;
            const uint8_t b = in[pos++];
            z |= (uint64_t)(b & 0x7f) << sh;
            sh += 7;
            if (!(b & 0x80)) break;
        }
        const int64_t d = (int64_t)(z >> 1) ^ -(int64_t)(z & 1);
        last += d;
        if (n >= cap) return -1;
        if (sample_bytes == 4) ((uint32_t*)x)[n] = (uint32_t)last; else ((uint16_t*)x)[n] = (uint16_t)last;
        ++n;
    }
    return n;
}
