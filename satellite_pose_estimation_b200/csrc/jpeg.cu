// Baseline JPEG (grayscale, Huffman, 8-bit) -> uint8 frame in HBM.
//
// Replaces `Image.open(img_path).convert('RGB')` at the head of the reference's data path (RV/datasets/speed.py:116 and
// :212; SA/src/data/speed/speed_dataset.py does the same): SPEED images are single-component baseline JPEG files, and
// `.convert('RGB')` of a mode-"L" image only replicates the channel -- which crop_resize_norm_kernel does anyway.  PIL
// decodes through libjpeg(-turbo) with its default DCT, JDCT_ISLOW; that inverse DCT is integer arithmetic with a fixed
// evaluation order (13-bit constants, 2 extra bits after the column pass), so the result is defined bit for bit and
// this file reproduces it bit for bit (tests/test_gpu_jpeg.py compares with PIL on every pixel).
//
// Work split.  The host walks the marker segments (a few hundred bytes per file), builds the canonical Huffman
// decoding tables and packs the entropy-coded scans of the batch into one pinned staging buffer -> one H2D copy of the
// COMPRESSED bytes (~10x fewer than the frames).  On the device a JPEG scan is sequential -- the bit position of a
// block is known only after its predecessors are decoded, and DC values are differences -- so one WARP owns one image:
//   lane 0      Huffman-decodes one 8x8 block at a time from a shared-memory window of the scan
//   all lanes   refill that window with coalesced 16-byte loads, zero the coefficient block
//   lanes 0-7   dequantise + column pass, then row pass of the ISLOW inverse DCT, 8-byte row stores into the frame
// Parallelism comes from the images in flight: a batch of 64 is 64 warps on 16 SMs' worth of CTAs and runs beside the
// tensor-core kernels of the other batches; an image set of thousands fills the machine.  Restart markers (DRI / RSTn),
// partial edge blocks, 16-bit quantisation tables and files with several DHT / DQT segments are handled; progressive
// (SOF2), arithmetic-coded, 12-bit and multi-component files are refused with SPE_ERR_INVALID -- the replicated-gray
// frame layout of this path cannot represent a colour image.
#include "spe_internal.h"
#include "profile.h"
#include "../../include/spe.h"

#include <string.h>

#include <atomic>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace spe {
int set_error(spe_ctx* ctx, int code, const std::string& msg);
}

namespace spe {
namespace {

constexpr int kLutBits = 10;
constexpr int kRing = 2048;       // bytes of scan window per warp (a power of two)
constexpr int kChunk = 512;       // refill granule: 32 lanes x 16 bytes

struct alignas(16) JpegTables {   // one per image, device memory
  uint16_t q[64];                 // quantisation table in the file's (zigzag) order
  uint16_t dc_lut[1 << kLutBits]; // (code length << 8) | symbol for codes of <= kLutBits bits, 0 = longer code
  uint16_t ac_lut[1 << kLutBits];
  int16_t ac_fast[1 << kLutBits]; // AC code AND its magnitude bits inside kLutBits bits: (value << 8) | (run << 4) |
                                  // total bits, 0 = take the two-step path (most coefficients are small: one lookup)
  int32_t dc_maxcode[18];         // [l] = largest code of length l (l = 1..16), -1 if none; [17] = sentinel
  int32_t ac_maxcode[18];
  int32_t dc_valoff[17];          // [l] = valptr[l] - mincode[l]
  int32_t ac_valoff[17];
  uint8_t dc_vals[256];
  uint8_t ac_vals[256];
};

struct JpegImage {                // one per image, device memory
  long long scan_off;             // byte offset of the entropy-coded data in the packed buffer (multiple of 16)
  int scan_len;
  int width, height;
  int restart;                    // MCUs (= blocks) per restart interval, 0 = none
};

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// libjpeg's range_limit table (jdmaster.c prepare_range_limit_table) as arithmetic: index = x & 1023 of the descaled
// value, centred on 128: [0,128) -> x + 128, [128,512) -> 255, [512,896) -> 0, [896,1024) -> x - 896
__device__ __forceinline__ uint32_t range_limit(int x) {
  const int i = x & 1023;
  return i < 128 ? i + 128 : (i < 512 ? 255 : (i < 896 ? 0 : i - 896));
}

// jidctint.c (jpeg_idct_islow): one 1-D pass over eight values, CONST_BITS = 13
__device__ __forceinline__ void idct_1d(const int (&in)[8], int (&out)[8], const int shift) {
  constexpr int F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299,
                F1_847 = 15137, F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * F0_541;
  int tmp2 = z1 + z3 * (-F1_847);
  int tmp3 = z1 + z2 * F0_765;
  z2 = in[0]; z3 = in[4];
  int tmp0 = (z2 + z3) << 13;
  int tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * F1_175;
  tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
  z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
  z3 += z5; z4 += z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  const int rnd = 1 << (shift - 1);
  out[0] = (tmp10 + tmp3 + rnd) >> shift; out[7] = (tmp10 - tmp3 + rnd) >> shift;
  out[1] = (tmp11 + tmp2 + rnd) >> shift; out[6] = (tmp11 - tmp2 + rnd) >> shift;
  out[2] = (tmp12 + tmp1 + rnd) >> shift; out[5] = (tmp12 - tmp1 + rnd) >> shift;
  out[3] = (tmp13 + tmp0 + rnd) >> shift; out[4] = (tmp13 - tmp0 + rnd) >> shift;
}

struct BitReader {                // lane 0 only
  const uint8_t* ring;
  long long pos;                  // absolute byte offset inside the scan of the next unread byte
  int len;
  uint64_t buf;
  int cnt;
  bool marker;                    // stopped in front of a marker (RSTn / EOI): zeros are fed until a restart clears it
  // top up to more than 32 valid bits: one call covers a whole symbol (code <= 16 bits + magnitude <= 11 bits).
  // Fast path: four scan bytes at once when none of them is 0xFF (no stuffing, no marker); else byte by byte.
  __device__ __forceinline__ void fill() {
    while (cnt <= 32) {
      if (!marker && pos + 4 <= len) {
        const uint32_t o = static_cast<uint32_t>(pos) & (kRing - 1);
        const uint32_t lo = *reinterpret_cast<const uint32_t*>(ring + (o & ~3u));
        const uint32_t hi = *reinterpret_cast<const uint32_t*>(ring + (((o & ~3u) + 4u) & (kRing - 1)));
        const uint32_t v = __funnelshift_r(lo, hi, (o & 3u) * 8u);          // bytes pos .. pos + 3, little endian
        const uint32_t nv = ~v;
        if (((nv - 0x01010101u) & ~nv & 0x80808080u) == 0u) {               // no byte of v is 0xFF
          buf |= static_cast<uint64_t>(__byte_perm(v, 0, 0x0123)) << (32 - cnt);
          cnt += 32;
          pos += 4;
          continue;
        }
      }
      uint32_t b = 0;
      if (!marker && pos < len) {
        b = ring[pos & (kRing - 1)];
        if (b == 0xFF) {
          const uint32_t b2 = pos + 1 < len ? ring[(pos + 1) & (kRing - 1)] : 0xD9u;
          if (b2 == 0) pos += 2;                  // stuffed zero
          else { marker = true; b = 0; }
        } else {
          ++pos;
        }
      }
      buf |= static_cast<uint64_t>(b) << (56 - cnt);
      cnt += 8;
    }
  }
  __device__ __forceinline__ uint32_t peek(int n) const { return static_cast<uint32_t>(buf >> (64 - n)); }
  __device__ __forceinline__ void skip(int n) { buf <<= n; cnt -= n; }
  __device__ __forceinline__ int receive_extend(int s) {   // F.2.2.1 EXTEND
    const int v = static_cast<int>(peek(s));
    skip(s);
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
  }
};

__device__ __forceinline__ int decode_symbol(BitReader& br, const uint16_t* lut, const int32_t* maxcode, const int32_t* valoff,
                                             const uint8_t* vals) {
  br.fill();
  const uint32_t e = lut[br.peek(kLutBits)];
  if (e != 0) {
    br.skip(static_cast<int>(e >> 8));
    return static_cast<int>(e & 0xFF);
  }
  int l = kLutBits + 1;
  int code = static_cast<int>(br.peek(l));
  while (l <= 16 && code > maxcode[l]) {
    ++l;
    code = static_cast<int>(br.peek(l));
  }
  if (l > 16) { br.skip(16); return 0; }          // corrupt data: libjpeg warns and uses 0
  br.skip(l);
  return vals[(code + valoff[l]) & 0xFF];
}

// Two warps per image: the entropy decoder (a scan is a sequential bit stream: one lane walks it, all 32 refill the
// window) and the inverse DCT + store warp, handing blocks over through a double-buffered coefficient array and one
// named barrier per block -- block n + 1 is Huffman-decoded while block n is transformed and stored.  (One warp doing both
// in turn spent a quarter of every block's time in the transform with the decoding lane idle.)
constexpr int kWarpsPerCta = 4;   // images per CTA (two warps each)
struct WarpSmem {
  JpegTables t;
  __align__(16) uint8_t ring[kRing];
  int coef[2][64];                // dequantised coefficients, natural order (double buffer: decoder -> transform warp)
  int last_k[2];                  // index of the last non-zero coefficient of the block in coef[i] (0: DC only)
  int ws[64];                     // after the column pass
};
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(kWarpsPerCta * 64)
jpeg_decode_kernel(const uint8_t* __restrict__ packed, const JpegImage* __restrict__ imgs, const JpegTables* __restrict__ tabs,
                   int B, uint8_t* __restrict__ frames, long long pitch, long long frame_stride, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp >> 1;                       // image of this CTA
  const bool is_decoder = (warp & 1) == 0;
  const int b = blockIdx.x * kWarpsPerCta + slot;
  if (b >= B) return;                               // both warps of the slot leave: no barrier is ever entered
  WarpSmem& S = reinterpret_cast<WarpSmem*>(smem_raw)[slot];
  const int bar_id = 1 + slot;
  if (is_decoder) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tabs + b);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&S.t);
    for (int i = lane; i < static_cast<int>(sizeof(JpegTables) / 4); i += 32) dst[i] = src[i];
  }
  const JpegImage im = imgs[b];
  const uint8_t* scan = packed + im.scan_off;
  const int bw = (im.width + 7) >> 3, bh = (im.height + 7) >> 3;
  uint8_t* frame = frames + static_cast<long long>(b) * frame_stride;
  const bool wide_store = (pitch % 8 == 0) && (reinterpret_cast<uintptr_t>(frame) % 8 == 0);

  if (!is_decoder) {
    // ---- transform warp: block n after the n-th barrier, from coef[n & 1]
    int n = 0;
    for (int by = 0; by < bh; ++by) {
      for (int bx = 0; bx < bw; ++bx, ++n) {
        pair_barrier(bar_id);
        const int* coef = S.coef[n & 1];
        const int last_k = S.last_k[n & 1];
        const int x0 = bx * 8, y0 = by * 8;
        // inverse DCT (jpeg_idct_islow); a DC-only block is its shortcut form, which the full formula reproduces
        if (last_k == 0) {
          if (lane < 8 && y0 + lane < im.height) {
            const uint32_t v = range_limit(((coef[0] << 2) + 16) >> 5);
            uint8_t* o = frame + static_cast<long long>(y0 + lane) * pitch + x0;
            if (wide_store && x0 + 8 <= im.width) {
              const uint32_t w = v * 0x01010101u;
              *reinterpret_cast<uint2*>(o) = make_uint2(w, w);
            } else {
              for (int i = 0; i < 8 && x0 + i < im.width; ++i) o[i] = static_cast<uint8_t>(v);
            }
          }
        } else {
          if (lane < 8) {                            // column `lane`
            int in[8], out[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) in[i] = coef[i * 8 + lane];
            idct_1d(in, out, 13 - 2);
#pragma unroll
            for (int i = 0; i < 8; ++i) S.ws[i * 8 + lane] = out[i];
          }
          __syncwarp();
          if (lane < 8 && y0 + lane < im.height) {   // row `lane`
            int in[8], out[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) in[i] = S.ws[lane * 8 + i];
            idct_1d(in, out, 13 + 2 + 3);
            uint8_t* o = frame + static_cast<long long>(y0 + lane) * pitch + x0;
            if (wide_store && x0 + 8 <= im.width) {
              const uint32_t lo = range_limit(out[0]) | (range_limit(out[1]) << 8) | (range_limit(out[2]) << 16) | (range_limit(out[3]) << 24);
              const uint32_t hi = range_limit(out[4]) | (range_limit(out[5]) << 8) | (range_limit(out[6]) << 16) | (range_limit(out[7]) << 24);
              *reinterpret_cast<uint2*>(o) = make_uint2(lo, hi);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (x0 + i < im.width) o[i] = static_cast<uint8_t>(range_limit(out[i]));
            }
          }
          __syncwarp();                              // ws is rewritten by the next block's column pass
        }
      }
    }
    return;
  }

  // ---- decoder warp: block n into coef[n & 1], then the n-th barrier (the transform warp reaches the (n + 1)-th only
  // after it has finished block n, so coef[n & 1] is free again when block n + 2 is written)
  long long filled = 0;            // the window holds scan bytes [filled - kRing, filled) (those still unread)
  BitReader br{S.ring, 0, im.scan_len, 0, 0, false};
  long long pos = 0;
  int pred = 0, until_restart = im.restart, bad = 0;
  __syncwarp();

  int nblk = 0;
  for (int by = 0; by < bh; ++by) {
    for (int bx = 0; bx < bw; ++bx, ++nblk) {
      int* coef = S.coef[nblk & 1];
      // ---- keep at least 1024 unread bytes in the window (one block consumes < 512 incl. stuffing)
      while (filled - pos < 1024 + kChunk && filled < im.scan_len + kChunk) {
        // the packed buffer is padded with 2 * kChunk zero bytes behind every scan: the over-read is harmless
        const uint4 v = *reinterpret_cast<const uint4*>(scan + filled + lane * 16);
        *reinterpret_cast<uint4*>(S.ring + ((filled + lane * 16) & (kRing - 1))) = v;
        filled += kChunk;
      }
      coef[lane] = 0;
      coef[lane + 32] = 0;
      __syncwarp();
      int last_k = 0;
      if (lane == 0) {
        br.pos = pos;
        if (im.restart > 0) {
          if (until_restart == 0) {
            // byte-align, step over the RSTn marker the reader stopped at, reset the DC prediction (F.2.2.4 / E.2.4)
            // (the reader never reads past a marker, so `pos` sits on it whether or not it has been seen yet)
            br.buf = 0; br.cnt = 0;
            if (br.pos + 1 < br.len && S.ring[br.pos & (kRing - 1)] == 0xFF &&
                (S.ring[(br.pos + 1) & (kRing - 1)] & 0xF8) == 0xD0)
              br.pos += 2;
            br.marker = false;
            pred = 0;
            until_restart = im.restart;
          }
          --until_restart;
        }
        const int t = decode_symbol(br, S.t.dc_lut, S.t.dc_maxcode, S.t.dc_valoff, S.t.dc_vals);
        if (t > 0) pred += br.receive_extend(t & 15);
        coef[0] = pred * static_cast<int>(S.t.q[0]);
        int k = 1;
        while (k < 64) {
          br.fill();
          const int f = S.t.ac_fast[br.peek(kLutBits)];
          if (f != 0) {                            // run, size and magnitude from one lookup
            k += (f >> 4) & 15;
            if (k > 63) { bad = 1; break; }
            br.skip(f & 15);
            coef[c_zigzag[k]] = (f >> 8) * static_cast<int>(S.t.q[k]);
            last_k = k;
            ++k;
            continue;
          }
          const int rs = decode_symbol(br, S.t.ac_lut, S.t.ac_maxcode, S.t.ac_valoff, S.t.ac_vals);
          const int r = rs >> 4, s = rs & 15;
          if (s == 0) {
            if (r != 15) break;                    // EOB
            k += 16;
            continue;
          }
          k += r;
          if (k > 63) { bad = 1; break; }
          coef[c_zigzag[k]] = br.receive_extend(s) * static_cast<int>(S.t.q[k]);
          last_k = k;
          ++k;
        }
        pos = br.pos;
        S.last_k[nblk & 1] = last_k;
      }
      pos = __shfl_sync(0xffffffffu, pos, 0);
      __syncwarp();
      pair_barrier(bar_id);                          // block nblk handed over
    }
  }
  if (lane == 0 && status != nullptr) status[b] = bad;
}

// ------------------------------------------------------------------------------------------------------------------
// host: marker segments -> tables
// ------------------------------------------------------------------------------------------------------------------
struct Huff {
  uint8_t counts[17] = {0};
  uint8_t vals[256] = {0};
  bool set = false;
};

void build_huff(const Huff& h, uint16_t* lut, int32_t* maxcode, int32_t* valoff, uint8_t* vals) {
  memset(lut, 0, sizeof(uint16_t) << kLutBits);
  memcpy(vals, h.vals, 256);
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    const int n = h.counts[l];
    valoff[l] = k - code;                           // valptr[l] - mincode[l]
    for (int i = 0; i < n; ++i, ++k, ++code) {
      if (l <= kLutBits) {
        const int lo = code << (kLutBits - l);
        for (int j = 0; j < (1 << (kLutBits - l)); ++j) lut[lo + j] = static_cast<uint16_t>((l << 8) | h.vals[k]);
      }
    }
    maxcode[l] = n ? code - 1 : -1;
    code <<= 1;
  }
  maxcode[0] = -1;
  maxcode[17] = 0x7fffffff;
  valoff[0] = 0;
}

struct Parsed {
  int width = 0, height = 0, restart = 0;
  size_t scan_begin = 0, scan_end = 0;
  JpegTables t;
};

inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

// "" on success
std::string parse_jpeg(const uint8_t* d, size_t n, Parsed* out) {
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return "not a JPEG file (no SOI marker)";
  uint16_t qt[4][64];
  bool qset[4] = {false, false, false, false};
  Huff dc[4], ac[4];
  int comp_tq = -1, td = -1, ta = -1;
  bool have_sof = false;
  size_t p = 2;
  while (p + 4 <= n) {
    if (d[p] != 0xFF) return "corrupt JPEG: marker expected";
    while (p < n && d[p] == 0xFF) ++p;               // fill bytes
    if (p >= n) break;
    const int m = d[p++];
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (m == 0xD9) break;
    if (p + 2 > n) return "corrupt JPEG: truncated segment";
    const size_t len = static_cast<size_t>(be16(d + p));
    if (len < 2 || p + len > n) return "corrupt JPEG: bad segment length";
    const uint8_t* s = d + p + 2;
    const size_t sl = len - 2;
    if (m == 0xDB) {                                 // DQT
      size_t i = 0;
      while (i < sl) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        ++i;
        if (tq > 3 || pq > 1 || i + (pq ? 128u : 64u) > sl) return "corrupt JPEG: bad DQT";
        for (int k = 0; k < 64; ++k) qt[tq][k] = static_cast<uint16_t>(pq ? be16(s + i + 2 * k) : s[i + k]);
        i += pq ? 128 : 64;
        qset[tq] = true;
      }
    } else if (m == 0xC4) {                          // DHT
      size_t i = 0;
      while (i < sl) {
        if (i + 17 > sl) return "corrupt JPEG: bad DHT";
        const int tc = s[i] >> 4, th = s[i] & 15;
        if (tc > 1 || th > 3) return "corrupt JPEG: bad DHT class / id";
        Huff& h = tc ? ac[th] : dc[th];
        int tot = 0;
        for (int l = 1; l <= 16; ++l) { h.counts[l] = s[i + l]; tot += s[i + l]; }
        i += 17;
        if (tot > 256 || i + tot > sl) return "corrupt JPEG: bad DHT";
        memset(h.vals, 0, 256);
        memcpy(h.vals, s + i, static_cast<size_t>(tot));
        i += static_cast<size_t>(tot);
        h.set = true;
      }
    } else if (m == 0xC0 || m == 0xC1) {             // SOF0 baseline / SOF1 extended sequential, Huffman
      if (sl < 6) return "corrupt JPEG: bad SOF";
      if (s[0] != 8) return "unsupported JPEG: only 8-bit samples";
      out->height = be16(s + 1);
      out->width = be16(s + 3);
      const int nf = s[5];
      if (nf != 1)
        return "unsupported JPEG: " + std::to_string(nf) + " components (this path carries single-channel frames; SPEED "
               "images are grayscale)";
      if (sl < 9) return "corrupt JPEG: bad SOF";
      comp_tq = s[8];
      have_sof = true;
    } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
      return m == 0xC2 ? "unsupported JPEG: progressive (SOF2)" : "unsupported JPEG: lossless / arithmetic-coded / hierarchical";
    } else if (m == 0xDD) {                          // DRI
      if (sl < 2) return "corrupt JPEG: bad DRI";
      out->restart = be16(s);
    } else if (m == 0xDA) {                          // SOS
      if (!have_sof) return "corrupt JPEG: SOS before SOF";
      if (sl < 6 || s[0] != 1) return "corrupt JPEG: bad SOS";
      td = s[2] >> 4; ta = s[2] & 15;
      if (td > 3 || ta > 3) return "corrupt JPEG: bad SOS table ids";
      size_t q = p + len;
      out->scan_begin = q;
      // the scan ends at the first marker that is neither a stuffed zero nor RSTn
      while (q + 1 < n) {
        const uint8_t* f = static_cast<const uint8_t*>(memchr(d + q, 0xFF, n - q - 1));
        if (!f) { q = n; break; }
        q = static_cast<size_t>(f - d);
        const int mm = d[q + 1];
        if (mm == 0 || (mm >= 0xD0 && mm <= 0xD7) || mm == 0xFF) { q += (mm == 0xFF) ? 1 : 2; continue; }
        break;
      }
      out->scan_end = q < n ? q : n;
      break;
    }
    p += len;
  }
  if (!have_sof || out->scan_end <= out->scan_begin) return "corrupt JPEG: no scan";
  if (out->width <= 0 || out->height <= 0) return "corrupt JPEG: empty image";
  if (comp_tq < 0 || comp_tq > 3 || !qset[comp_tq]) return "corrupt JPEG: quantisation table missing";
  if (!dc[td].set || !ac[ta].set) return "corrupt JPEG: Huffman table missing";
  memcpy(out->t.q, qt[comp_tq], sizeof(out->t.q));
  build_huff(dc[td], out->t.dc_lut, out->t.dc_maxcode, out->t.dc_valoff, out->t.dc_vals);
  build_huff(ac[ta], out->t.ac_lut, out->t.ac_maxcode, out->t.ac_valoff, out->t.ac_vals);
  for (int i = 0; i < (1 << kLutBits); ++i) {
    const int e = out->t.ac_lut[i];
    out->t.ac_fast[i] = 0;
    const int len = e >> 8, run = (e >> 4) & 15, mag = e & 15;
    if (e == 0 || mag == 0 || len + mag > kLutBits) continue;
    int v = ((i << len) & ((1 << kLutBits) - 1)) >> (kLutBits - mag);      // the magnitude bits behind the code
    if (v < (1 << (mag - 1))) v += -(1 << mag) + 1;                        // EXTEND
    if (v >= -128 && v <= 127) out->t.ac_fast[i] = static_cast<int16_t>(v * 256 + run * 16 + len + mag);
  }
  return "";
}

// run f(i) for i in [0, n) on up to 16 host threads (header walks and scan copies of a batch are independent per file)
template <typename F>
void parallel_for(int n, F f) {
  const unsigned hw = std::thread::hardware_concurrency();
  const int nt = n < 16 ? 1 : static_cast<int>(hw == 0 ? 4 : (hw > 16 ? 16 : hw));
  if (nt <= 1) {
    for (int i = 0; i < n; ++i) f(i);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&]() {
      for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i);
    });
  for (auto& t : th) t.join();
}

// The compressed bytes travel through two fixed pinned windows (the host threads pack window w + 1 while window w is
// on the wire) into one device buffer that grows with the largest batch seen; a pinned buffer the size of a whole
// image set (gigabytes) would cost more to allocate than the set takes to decode.
constexpr size_t kStageWindow = size_t(96) << 20;
struct JpegState {
  uint8_t* win_h[2] = {nullptr, nullptr};   // pinned windows
  size_t win_cap = 0;
  cudaEvent_t win_free[2] = {nullptr, nullptr};   // fires when the copy out of the window has completed
  bool win_busy[2] = {false, false};
  uint8_t* stage_d = nullptr;               // device: [tables | image records | packed scans]
  size_t cap_d = 0;
  int* status_d = nullptr;
  int status_cap = 0;
};
std::map<spe_ctx*, JpegState*> g_jpeg;

}  // namespace

void jpeg_release(spe_ctx* ctx) {
  auto it = g_jpeg.find(ctx);
  if (it == g_jpeg.end()) return;
  JpegState* s = it->second;
  for (int w = 0; w < 2; ++w) {
    if (s->win_busy[w]) cudaEventSynchronize(s->win_free[w]);
    if (s->win_h[w]) cudaFreeHost(s->win_h[w]);
    if (s->win_free[w]) cudaEventDestroy(s->win_free[w]);
  }
  if (s->stage_d) cudaFree(s->stage_d);
  if (s->status_d) cudaFree(s->status_d);
  delete s;
  g_jpeg.erase(it);
}

}  // namespace spe

using namespace spe;

extern "C" {

int spe_jpeg_info(const uint8_t* file_host, long long size, int* width, int* height) {
  if (!file_host || size <= 0) return set_error(nullptr, SPE_ERR_INVALID, "spe_jpeg_info: null / empty file");
  Parsed ps;
  const std::string e = parse_jpeg(file_host, static_cast<size_t>(size), &ps);
  if (!e.empty()) return set_error(nullptr, SPE_ERR_INVALID, "spe_jpeg_info: " + e);
  if (width) *width = ps.width;
  if (height) *height = ps.height;
  return SPE_OK;
}

int spe_jpeg_decode_batch(spe_ctx* ctx, const uint8_t* const* files_host, const long long* sizes, int B, uint8_t* frames_dev,
                          int H, int W, long long pitch, long long frame_stride, void* stream) {
  if (!ctx) return set_error(nullptr, SPE_ERR_INVALID, "spe_jpeg_decode_batch: null ctx");
  if (!files_host || !sizes || !frames_dev || B <= 0)
    return set_error(ctx, SPE_ERR_INVALID, "spe_jpeg_decode_batch: null argument / empty batch");
  if (pitch < W || frame_stride < static_cast<long long>(H) * pitch)
    return set_error(ctx, SPE_ERR_INVALID, "spe_jpeg_decode_batch: pitch / frame stride smaller than the frame");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  JpegState*& S = g_jpeg[ctx];
  if (!S) S = new JpegState();
  std::vector<Parsed> ps(static_cast<size_t>(B));
  const size_t tab_bytes = sizeof(JpegTables) * static_cast<size_t>(B);
  const size_t img_bytes = ((sizeof(JpegImage) * static_cast<size_t>(B)) + 15) / 16 * 16;
  size_t total = tab_bytes + img_bytes;
  std::vector<size_t> offs(static_cast<size_t>(B));
  std::vector<std::string> errs(static_cast<size_t>(B));
  parallel_for(B, [&](int i) {
    if (!files_host[i] || sizes[i] <= 0) { errs[i] = "null / empty"; return; }
    errs[i] = parse_jpeg(files_host[i], static_cast<size_t>(sizes[i]), &ps[i]);
    if (errs[i].empty() && (ps[i].width != W || ps[i].height != H))
      errs[i] = "is " + std::to_string(ps[i].width) + "x" + std::to_string(ps[i].height) + ", the frame buffer " +
                std::to_string(W) + "x" + std::to_string(H);
  });
  for (int i = 0; i < B; ++i) {
    if (!errs[i].empty())
      return set_error(ctx, SPE_ERR_INVALID, "spe_jpeg_decode_batch: file " + std::to_string(i) + ": " + errs[i]);
    offs[i] = total;
    total += (ps[i].scan_end - ps[i].scan_begin + 15) / 16 * 16 + 2 * kChunk;     // zero padding behind every scan
  }
  // ---- buffers: device buffer for everything, two pinned windows at least as large as the header block / any one scan
  size_t need_win = tab_bytes + img_bytes;
  for (int i = 0; i < B; ++i) {
    const size_t sz = (i + 1 < B ? offs[i + 1] : total) - offs[i];
    if (sz > need_win) need_win = sz;
  }
  if (need_win < kStageWindow) need_win = kStageWindow;
  if (need_win > S->win_cap) {
    for (int w = 0; w < 2; ++w) {
      if (S->win_busy[w]) { cudaEventSynchronize(S->win_free[w]); S->win_busy[w] = false; }
      if (S->win_h[w]) cudaFreeHost(S->win_h[w]);
      S->win_h[w] = nullptr;
      if (cudaMallocHost(reinterpret_cast<void**>(&S->win_h[w]), need_win) != cudaSuccess) {
        cudaGetLastError();
        S->win_cap = 0;
        return set_error(ctx, SPE_ERR_CUDA, "spe_jpeg_decode_batch: out of pinned memory for the staging windows");
      }
      if (!S->win_free[w] && cudaEventCreateWithFlags(&S->win_free[w], cudaEventDisableTiming) != cudaSuccess)
        return set_error(ctx, SPE_ERR_CUDA, "spe_jpeg_decode_batch: cudaEventCreate failed");
    }
    S->win_cap = need_win;
  }
  if (total > S->cap_d) {
    // the previous batch's kernel may still read the old buffer: cudaFree waits for the device
    if (S->stage_d) cudaFree(S->stage_d);
    S->stage_d = nullptr; S->cap_d = 0;
    const size_t want = total + total / 4;
    if (cudaMalloc(reinterpret_cast<void**>(&S->stage_d), want) != cudaSuccess) {
      cudaGetLastError();
      return set_error(ctx, SPE_ERR_CUDA, "spe_jpeg_decode_batch: out of device memory for the compressed scans");
    }
    S->cap_d = want;
  }
  if (B > S->status_cap) {
    if (S->status_d) cudaFree(S->status_d);
    S->status_d = nullptr; S->status_cap = 0;
    if (cudaMalloc(reinterpret_cast<void**>(&S->status_d), sizeof(int) * B) != cudaSuccess) {
      cudaGetLastError();
      return set_error(ctx, SPE_ERR_CUDA, "spe_jpeg_decode_batch: out of memory");
    }
    S->status_cap = B;
  }
  cudaError_t e = cudaSuccess;
  int w = 0;
  auto window = [&]() -> uint8_t* {          // next window, once the copy that last used it has completed
    w ^= 1;
    if (S->win_busy[w]) { cudaEventSynchronize(S->win_free[w]); S->win_busy[w] = false; }
    return S->win_h[w];
  };
  auto ship = [&](size_t dev_off, size_t bytes) {
    if (e != cudaSuccess) return;
    e = cudaMemcpyAsync(S->stage_d + dev_off, S->win_h[w], bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaEventRecord(S->win_free[w], st);
    S->win_busy[w] = true;
  };
  // ---- header block: tables + image records
  {
    uint8_t* h = window();
    JpegTables* th = reinterpret_cast<JpegTables*>(h);
    JpegImage* ih = reinterpret_cast<JpegImage*>(h + tab_bytes);
    parallel_for(B, [&](int i) {
      th[i] = ps[i].t;
      ih[i].scan_off = static_cast<long long>(offs[i]);
      ih[i].scan_len = static_cast<int>(ps[i].scan_end - ps[i].scan_begin);
      ih[i].width = ps[i].width;
      ih[i].height = ps[i].height;
      ih[i].restart = ps[i].restart;
    });
    ship(0, tab_bytes + img_bytes);
  }
  // ---- scans, a window's worth of whole files at a time
  for (int i0 = 0; i0 < B;) {
    int i1 = i0;
    const size_t base = offs[i0];
    while (i1 < B && (i1 + 1 < B ? offs[i1 + 1] : total) - base <= S->win_cap) ++i1;
    uint8_t* h = window();
    parallel_for(i1 - i0, [&](int k) {
      const int i = i0 + k;
      const size_t len = ps[i].scan_end - ps[i].scan_begin;
      memcpy(h + (offs[i] - base), files_host[i] + ps[i].scan_begin, len);
      const size_t padded = (i + 1 < B ? offs[i + 1] : total) - offs[i];
      memset(h + (offs[i] - base) + len, 0, padded - len);
    });
    ship(base, (i1 < B ? offs[i1] : total) - base);
    i0 = i1;
  }
  if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("spe_jpeg_decode_batch: ") + cudaGetErrorString(e));
  static bool attr = false;
  const size_t smem = sizeof(WarpSmem) * kWarpsPerCta;
  if (!attr) {
    e = cudaFuncSetAttribute(jpeg_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("spe_jpeg_decode_batch: ") + cudaGetErrorString(e));
    attr = true;
  }
  {
    ProfScope psc(kFamCrop, st);
    jpeg_decode_kernel<<<(B + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 64, smem, st>>>(
        S->stage_d, reinterpret_cast<const JpegImage*>(S->stage_d + tab_bytes), reinterpret_cast<const JpegTables*>(S->stage_d), B,
        frames_dev, pitch, frame_stride, S->status_d);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(ctx, SPE_ERR_CUDA, std::string("spe_jpeg_decode_batch: ") + cudaGetErrorString(e));
  return SPE_OK;
}

}  // extern "C"
