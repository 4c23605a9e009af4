// Baseline JPEG decode on the device (sm_100a): header parser (host), entropy decode kernel, IDCT / upsample / colour kernel.
//
// Replaces tf.io.read_file -> tf.image.decode_jpeg(channels=3) (dataset/dataset.py:24-28), i.e. libjpeg-turbo with its
// defaults: JDCT_ISLOW integer inverse DCT (jidctint.c), fancy ("triangle") chroma upsampling (jdsample.c
// h2v1_fancy_upsample / h2v2_fancy_upsample) and the 16-bit fixed-point YCbCr -> RGB of jdcolor.c.  Restated from the
// published algorithms (ITU-T T.81 for the stream syntax and the Huffman procedure of Annex F; the libjpeg documentation for
// the integer pipelines); the integer IDCT / upsampling / colour arithmetic is shared with the JPEG-quality emulation of
// preprocess.cu (jpeg_math.cuh) and is pinned bit for bit against libjpeg-turbo through Pillow (tests/test_jpeg_decode_gpu.py,
// tests/golden/jpeg_files.npz).  oracle/jpeg_decode.py is the checker, never linked here.
//
// Kernel 1 (jpeg_entropy_kernel): one warp per image.  The warp builds the four decoding tables of the image in shared
// memory (an 11-bit lookahead table that resolves code + magnitude bits in one lookup when both fit in the window, and the
// maxcode / valptr arrays of T.81 F.2.2.3 for longer codes); lane 0 then walks the
// entropy-coded segment (64-bit bit buffer refilled a 32-bit word at a time when the word holds no 0xFF, byte by byte around
// stuffed zeros, restart markers and the end of the scan) and fills one 8x8 block of coefficients in shared memory; the
// whole warp writes the block (128 bytes, zeros included) with one coalesced store, so the workspace needs no memset.
// Coefficients are stored column-major inside a block: the column pass of the IDCT then reads 16 contiguous bytes per lane.
// The stream of one image is inherently sequential (no restart markers in the files of this path); the batch supplies the
// parallelism -- 1024 images = 1024 warps, ~7 per SM, each a latency-bound dependent chain.
//
// Kernel 2 (jpeg_pixels_kernel): one CTA per (band of MCU rows, image).  8 lanes per block: dequantise, IDCT columns,
// 8-lane transpose, IDCT rows (+128, clamp) into u8 component planes in shared memory (chroma with one block row of halo
// above and below the band when it is vertically subsampled), then upsample + colour-convert + store interleaved RGB.
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "jpeg_math.cuh"

namespace vip {
namespace {

constexpr int kLook = 11;                 // lookahead bits of the AC tables' fast path (2 x 2048 x 4 B per image)
constexpr int kLookDC = 9;                // ... of the DC tables (one symbol in 64 is a DC difference)
constexpr int kPixThreads = 256;
constexpr int kPixWarps = kPixThreads / 32;
constexpr int kPlaneBudget = 96 * 1024;   // bytes of component planes per CTA (2 CTAs per SM)

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t h_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                              41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                              30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Block geometry of one image (T.81 A.1.1, A.2): the scan is interleaved, MCU = hs*vs blocks of every component.
struct Geo {
  int hmax, vmax, mcux, mcuy;
  int nbx[3], nby[3];   // blocks per row / column of each component (padded to whole MCUs)
  int cboff[3];         // first block of each component inside the image's coefficient slice
  int nblocks;
  int mode;             // 0 grey, 1 4:4:4, 2 4:2:2 (h2v1), 3 4:2:0 (h2v2)
  int band_bytes_per_row, band_halo_bytes;   // shared-memory plane bytes per MCU row of a band / for the chroma halo
};

__host__ __device__ inline Geo geometry(const vip_jpeg_desc& d) {
  Geo g;
  g.hmax = d.hs[0];
  g.vmax = d.vs[0];
  g.mcux = (d.width + 8 * g.hmax - 1) / (8 * g.hmax);
  g.mcuy = (d.height + 8 * g.vmax - 1) / (8 * g.vmax);
  int off = 0;
  g.band_bytes_per_row = 0;
  g.band_halo_bytes = 0;
  for (int c = 0; c < 3; ++c) {
    const bool live = c < d.ncomp;
    g.nbx[c] = live ? g.mcux * d.hs[c] : 0;
    g.nby[c] = live ? g.mcuy * d.vs[c] : 0;
    g.cboff[c] = off;
    off += g.nbx[c] * g.nby[c];
    if (live) {
      g.band_bytes_per_row += d.vs[c] * 8 * g.nbx[c] * 8;
      if (d.vs[c] < g.vmax) g.band_halo_bytes += 2 * 8 * g.nbx[c] * 8;
    }
  }
  g.nblocks = off;
  g.mode = d.ncomp == 1 ? 0 : (g.hmax == 1 ? 1 : (g.vmax == 1 ? 2 : 3));
  return g;
}

__host__ __device__ inline int band_rows(const Geo& g) {
  const int r = (kPlaneBudget - g.band_halo_bytes) / g.band_bytes_per_row;
  return r < 1 ? 1 : (r > g.mcuy ? g.mcuy : r);
}

// ---- kernel 1: entropy decode -----------------------------------------------------------------------------------------
// Bit reader of the one thread that walks a stream.  The next bits of the stream sit at the TOP of a 64-bit register (n of
// them valid, zeros below), so the 32-bit decoding window is simply its high word; consuming k bits is one 64-bit shift.
// Positions are 32-bit offsets from the start of the entropy-coded segment (short dependent chains: the thread is bound by
// instruction latency, not by throughput).
struct BitReader {
  const uint8_t* base;      // first byte of the entropy-coded segment
  unsigned pos, end;        // next byte, segment length
  unsigned mis;             // (address of base) & 3: word loads need pos + mis to be a multiple of 4
  unsigned long long acc;   // MSB-aligned bit buffer
  int n;                    // valid bits in acc
  bool marker;              // a marker (or the end of the data) stops the stream: zeros are fed from there on
  int fake;                 // zero bits fed past a marker / the end of the data and still inside acc or already consumed
  bool have_next;           // next_w holds the (aligned) word at pos, requested one refill earlier: the load latency of
  unsigned next_w;          // the stream is off the dependent chain (L1 is small next to the tables of ~10 images)
};

__device__ __forceinline__ void br_byte(BitReader& br) {
  unsigned b = 0;
  br.have_next = false;
  if (!br.marker) {
    if (br.pos < br.end) {
      b = __ldg(br.base + br.pos++);
      if (b == 0xFFu) {
        const unsigned b2 = br.pos < br.end ? __ldg(br.base + br.pos) : 0xD9u;
        if (b2 == 0u) {
          ++br.pos;               // stuffed zero (T.81 B.1.1.5)
        } else {
          br.marker = true;       // leave pos on the 0xFF of the marker
          --br.pos;
          b = 0;
        }
      }
    } else {
      br.marker = true;
    }
  }
  if (br.marker) br.fake += 8;
  br.acc |= (unsigned long long)b << (56 - br.n);
  br.n += 8;
}

// more than 32 valid bits afterwards
__device__ __forceinline__ void br_refill(BitReader& br) {
  while (br.n <= 32) {
    if (!br.marker && ((br.pos + br.mis) & 3u) == 0u && br.pos + 4u <= br.end) {
      const unsigned w = br.have_next ? br.next_w : __ldg(reinterpret_cast<const unsigned*>(br.base + br.pos));
      if ((((~w) - 0x01010101u) & w & 0x80808080u) == 0u) {   // no byte of w is 0xFF
        br.acc |= (unsigned long long)__byte_perm(w, 0u, 0x0123) << (32 - br.n);
        br.n += 32;
        br.pos += 4u;
        br.have_next = br.pos + 4u <= br.end;
        if (br.have_next) br.next_w = __ldg(reinterpret_cast<const unsigned*>(br.base + br.pos));   // used ~3 symbols later
        continue;
      }
    }
    br_byte(br);
  }
}
__device__ __forceinline__ unsigned br_window(const BitReader& br) { return (unsigned)(br.acc >> 32); }
__device__ __forceinline__ void br_consume(BitReader& br, int k) {
  br.acc <<= k;
  br.n -= k;
}

// Decoding tables of one image.  `fast_*` is indexed by the next kLook (AC) / kLookDC bits of the stream and resolves, in ONE lookup, the
// Huffman code AND the magnitude bits that follow it whenever both fit in the window (the common case: short codes of
// small coefficients).  The single thread that walks a stream is bound by the latency of its dependent instruction chain,
// so the instructions per symbol are what counts:
//   bit 15 set    complete symbol: bits 0-4 = code length + magnitude bits, 5-8 = zero run, bit 9 = no magnitude (EOB / ZRL
//                 / DC difference 0), bits 16-31 = the EXTENDed value (T.81 F.2.2.1)
//   bit 15 clear  nonzero: code of <= kLook bits whose magnitude bits spill over the window: (code length << 8) | symbol
//                 zero: code longer than kLook bits -> maxcode / valptr search (T.81 F.2.2.3)
struct HuffTables {
  unsigned fast_dc[2][1 << kLookDC];
  unsigned fast_ac[2][1 << kLook];
  int maxcode[4][17];                   // largest code of each length, -1 if none
  int valoff[4][17];                    // index of the first symbol of that length minus its code
  uint8_t vals[4][256];
};

// T.81 F.2.2.1 EXTEND of the s-bit value v
__device__ __forceinline__ int huff_extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// One symbol from the 32-bit window w (the next 32 bits of the stream): returns the bits consumed; run / has_val / val out.
// t = table index into maxcode / valoff / vals (0, 1 DC; 2, 3 AC); kBits = lookahead of that table's fast path.
template <int kBits>
__device__ __forceinline__ int huff_symbol(unsigned w, const HuffTables& T, int t, int& run, bool& has_val, int& val, bool& bad) {
  const unsigned e = kBits == kLookDC ? T.fast_dc[t & 1][w >> (32 - kBits)] : T.fast_ac[t & 1][w >> (32 - kBits)];
  if (e & 0x8000u) {
    run = (int)((e >> 5) & 15u);
    has_val = (e & 0x200u) == 0u;
    val = (int)e >> 16;
    return (int)(e & 31u);
  }
  int len, rs;
  if (e != 0u) {
    len = (int)(e >> 8);
    rs = (int)(e & 255u);
  } else {
    len = 0;
    rs = 0;
#pragma unroll 1
    for (int l = kBits + 1; l <= 16; ++l) {
      const int code = (int)(w >> (32 - l));
      if (code <= T.maxcode[t][l]) {
        len = l;
        rs = T.vals[t][(T.valoff[t][l] + code) & 255];
        break;
      }
    }
    if (len == 0) {
      bad = true;
      len = 16;
    }
  }
  run = rs >> 4;
  const int s = rs & 15;
  has_val = s != 0;
  val = 0;
  if (s) val = huff_extend((int)((w << len) >> (32 - s)), s);   // len + s <= 31: inside the window
  return len + s;
}

__global__ void __launch_bounds__(32) jpeg_entropy_kernel(const uint8_t* __restrict__ data, const vip_jpeg_desc* __restrict__ descs,
                                                        int16_t* __restrict__ coef, int32_t* __restrict__ err) {
  pdl_trigger();
  pdl_wait();
  __shared__ HuffTables T;
  __shared__ __align__(16) int16_t blk[64];
  __shared__ uint8_t zzT[64];
  const int lane = threadIdx.x;
  const vip_jpeg_desc& d = descs[blockIdx.x];
  if (d.status != VIP_JPEG_OK) {
    if (err != nullptr && lane == 0) err[blockIdx.x] = 0;
    return;
  }
  const Geo g = geometry(d);
  // ---- tables
  for (int i = lane; i < 4 * 256 / 4; i += 32)
    reinterpret_cast<unsigned*>(&T.vals[0][0])[i] = reinterpret_cast<const unsigned*>(&d.huff_vals[0][0])[i];
  for (int i = lane; i < 64; i += 32) {
    const int nat = c_zigzag[i];
    zzT[i] = (uint8_t)((nat & 7) * 8 + (nat >> 3));
  }
  reinterpret_cast<unsigned*>(blk)[lane] = 0u;
  __syncwarp();
  if (lane < 4) {   // T.81 Annex C: codes in order of length; maxcode / valptr per length
    const int t = lane;
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      const int nb = d.huff_bits[t][l - 1];
      T.maxcode[t][l] = nb ? code + nb - 1 : -1;
      T.valoff[t][l] = k - code;
      code = (code + nb) << 1;
      k += nb;
    }
  }
  __syncwarp();
  // fast tables: every lane resolves the windows i = lane, lane + 32, ... of the four tables with the canonical search
  for (int i = lane; i < 2 * (1 << kLookDC) + 2 * (1 << kLook); i += 32) {
    const bool dc = i < 2 * (1 << kLookDC);
    const int bits = dc ? kLookDC : kLook;
    const int j = dc ? i : i - 2 * (1 << kLookDC);
    const int t = (dc ? 0 : 2) + (j >> bits);
    const unsigned win = (unsigned)(j & ((1 << bits) - 1));
    unsigned e = 0u;
    for (int l = 1; l <= bits; ++l) {
      const int code = (int)(win >> (bits - l));
      if (code <= T.maxcode[t][l]) {
        const int rs = T.vals[t][(T.valoff[t][l] + code) & 255];
        const int sz = rs & 15;
        if (l + sz <= bits) {
          int v = 0;
          if (sz) v = huff_extend((int)((win >> (bits - l - sz)) & ((1u << sz) - 1u)), sz);
          e = ((unsigned)(v & 0xFFFF) << 16) | 0x8000u | (sz ? 0u : 0x200u) | ((unsigned)(rs >> 4) << 5) | (unsigned)(l + sz);
        } else {
          e = ((unsigned)l << 8) | (unsigned)rs;
        }
        break;
      }
    }
    if (dc) T.fast_dc[t & 1][win] = e;
    else T.fast_ac[t & 1][win] = e;
  }
  __syncwarp();

  BitReader br;
  br.base = data + d.file_offset + d.scan_offset;
  br.pos = 0u;
  br.end = (unsigned)d.scan_bytes;
  br.mis = (unsigned)(reinterpret_cast<uintptr_t>(br.base) & 3u);
  br.acc = 0;
  br.n = 0;
  br.marker = false;
  br.have_next = false;
  br.next_w = 0u;
  br.fake = 0;
  int pred0 = 0, pred1 = 0, pred2 = 0;
  bool bad = false;
  int until_restart = d.restart_interval;
  int16_t* out = coef + d.coef_offset * 64;
  const int ncomp = d.ncomp, ri = d.restart_interval;

  for (int my = 0; my < g.mcuy; ++my) {
    for (int mx = 0; mx < g.mcux; ++mx) {
      if (lane == 0 && ri > 0) {
        if (until_restart == 0) {
          // T.81 F.2.2.5 / E.2.4: byte-align, expect RSTm, reset the predictors
          if (br.fake > br.n) bad = true;          // the interval consumed bits that were never in the file
          br.fake = 0;
          br.n = 0;
          br.acc = 0;
          if (!br.marker) br_byte(br);             // must run into the marker at once
          br.n = 0;
          br.acc = 0;
          if (br.marker && br.pos + 2u <= br.end && br.base[br.pos] == 0xFF && (br.base[br.pos + 1] & 0xF8) == 0xD0) {
            br.pos += 2u;
            br.marker = false;
            br.fake = 0;
          } else {
            bad = true;
          }
          pred0 = pred1 = pred2 = 0;
          until_restart = ri;
        }
        --until_restart;
      }
#pragma unroll 1
      for (int c = 0; c < ncomp; ++c) {
        const int hs = d.hs[c], vs = d.vs[c];
        const int td = d.td[c], ta = 2 + d.ta[c];
#pragma unroll 1
        for (int b = 0; b < hs * vs; ++b) {
          const int v = b / hs, h = b - v * hs;
          if (lane == 0 && !bad) {
            int run, val;
            bool has_val;
            br_refill(br);
            br_consume(br, huff_symbol<kLookDC>(br_window(br), T, td, run, has_val, val, bad));
            int pred = c == 0 ? pred0 : (c == 1 ? pred1 : pred2);
            pred += val;
            if (c == 0) pred0 = pred;
            else if (c == 1) pred1 = pred;
            else pred2 = pred;
            blk[0] = (int16_t)pred;
            int k = 1;
#pragma unroll 1
            while (k < 64) {
              br_refill(br);
              br_consume(br, huff_symbol<kLook>(br_window(br), T, ta, run, has_val, val, bad));
              if (has_val) {
                k += run;
                blk[zzT[k & 63]] = (int16_t)val;
                if (k > 63) bad = true;
                ++k;
              } else if (run == 15) {
                k += 16;
              } else {
                break;
              }
            }
          }
          __syncwarp();
          const size_t bi = (size_t)g.cboff[c] + (size_t)(my * vs + v) * g.nbx[c] + (mx * hs + h);
          reinterpret_cast<unsigned*>(out + bi * 64)[lane] = reinterpret_cast<const unsigned*>(blk)[lane];
          reinterpret_cast<unsigned*>(blk)[lane] = 0u;
          __syncwarp();
        }
      }
    }
  }
  // zero bits fed after the marker that ends the scan are legitimate read-ahead only while they stay unconsumed (libjpeg
  // warns "premature end of data segment" and pads; tf.image.decode_jpeg rejects such a file)
  if (err != nullptr && lane == 0) err[blockIdx.x] = (bad || br.fake > br.n) ? 1 : 0;
}

// ---- kernel 2: dequantise, IDCT, upsample, colour ------------------------------------------------------------------------
__device__ __forceinline__ void store_rgb_pair(uint8_t* p, const int (&px)[6], int count) {
  if (count == 2 && (reinterpret_cast<uintptr_t>(p) & 1) == 0) {
    unsigned short* q = reinterpret_cast<unsigned short*>(p);
    q[0] = (unsigned short)(px[0] | (px[1] << 8));
    q[1] = (unsigned short)(px[2] | (px[3] << 8));
    q[2] = (unsigned short)(px[4] | (px[5] << 8));
  } else {
    for (int k = 0; k < 3 * count; ++k) p[k] = (uint8_t)px[k];
  }
}

// jdcolor.c ycc_rgb_convert: R = y + ((91881 cr' + 32768) >> 16), G = y + ((-22554 cb' - 46802 cr' + 32768) >> 16),
// B = y + ((116130 cb' + 32768) >> 16) with cb' = cb - 128, cr' = cr - 128 (y and the -128 folded into one addend)
__device__ __forceinline__ void ycc_to_rgb(int y, int cb, int cr, int* rgb) {
  const int yb = (y << 16) + 32768;
  rgb[0] = __vimin_s32_relu(y + ((91881 * (cr - 128) + 32768) >> 16), 255);
  rgb[1] = __vimin_s32_relu((-22554 * cb - 46802 * cr + (yb + 128 * (22554 + 46802))) >> 16, 255);
  rgb[2] = __vimin_s32_relu(y + ((116130 * (cb - 128) + 32768) >> 16), 255);
}

__global__ void __launch_bounds__(kPixThreads) jpeg_pixels_kernel(const vip_jpeg_desc* __restrict__ descs,
                                                                  const int16_t* __restrict__ coef, uint8_t* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_qt[3][64];                       // transposed like the coefficients: [u * 8 + v]
  const vip_jpeg_desc& d = descs[blockIdx.y];
  if (d.status != VIP_JPEG_OK) return;
  const Geo g = geometry(d);
  const int R = band_rows(g);
  const int m0 = blockIdx.x * R;
  if (m0 >= g.mcuy) return;
  const int m1 = min(m0 + R, g.mcuy);
  const int tid = threadIdx.x;

  // block-row range [br0, br1) of each component held in shared memory, plane base offsets
  int br0[3], br1[3], pbase[3], pitch[3], nblk[3];
  int off = 0, total = 0;
  for (int c = 0; c < 3; ++c) {
    if (c < d.ncomp) {
      const bool halo = d.vs[c] < g.vmax;
      br0[c] = halo ? max(m0 - 1, 0) : m0 * d.vs[c];
      br1[c] = halo ? min(m1 + 1, g.mcuy) : m1 * d.vs[c];
      pitch[c] = g.nbx[c] * 8;
      pbase[c] = off;
      off += (br1[c] - br0[c]) * 8 * pitch[c];
      nblk[c] = (br1[c] - br0[c]) * g.nbx[c];
    } else {
      br0[c] = br1[c] = pbase[c] = pitch[c] = nblk[c] = 0;
    }
    total += nblk[c];
  }
  for (int i = tid; i < d.ncomp * 64; i += kPixThreads) {
    const int c = i >> 6, k = i & 63;   // k = u * 8 + v  <-  natural v * 8 + u
    s_qt[c][k] = d.qt[d.tq[c]][(k & 7) * 8 + (k >> 3)];
  }
  __syncthreads();

  // ---- IDCT: 8 lanes per block, 4 blocks per warp
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int b = lane >> 3, r = lane & 7;
    const int16_t* cimg = coef + d.coef_offset * 64;
    const int iters = (total + 3) >> 2;
    for (int it = warp; it < iters; it += kPixWarps) {
      int bi = it * 4 + b;
      const bool live = bi < total;
      bi = live ? bi : total - 1;
      int c = 0;
      if (bi >= nblk[0]) { bi -= nblk[0]; c = 1; }
      if (c == 1 && bi >= nblk[1]) { bi -= nblk[1]; c = 2; }
      const int by = bi / g.nbx[c], bx = bi - by * g.nbx[c];
      const int16_t* cb = cimg + ((size_t)g.cboff[c] + (size_t)(br0[c] + by) * g.nbx[c] + bx) * 64;
      const int4 raw = __ldg(reinterpret_cast<const int4*>(cb + r * 8));   // column u = r, v = 0..7
      const int4 q0 = *reinterpret_cast<const int4*>(&s_qt[c][r * 8]);
      const int4 q1 = *reinterpret_cast<const int4*>(&s_qt[c][r * 8 + 4]);
      int dd[8];
      dd[0] = (int)(short)(raw.x & 0xFFFF) * q0.x; dd[1] = (raw.x >> 16) * q0.y;
      dd[2] = (int)(short)(raw.y & 0xFFFF) * q0.z; dd[3] = (raw.y >> 16) * q0.w;
      dd[4] = (int)(short)(raw.z & 0xFFFF) * q1.x; dd[5] = (raw.z >> 16) * q1.y;
      dd[6] = (int)(short)(raw.w & 0xFFFF) * q1.z; dd[7] = (raw.w >> 16) * q1.w;
      idct8<true>(dd);                 // column pass (over v)
      transpose8_shfl(dd, r);          // lane r holds row y = r
      idct8<false>(dd);
      uint2 o;
      o.x = __vimin_s32_relu(dd[0], 255) | (__vimin_s32_relu(dd[1], 255) << 8) | (__vimin_s32_relu(dd[2], 255) << 16) |
            (__vimin_s32_relu(dd[3], 255) << 24);
      o.y = __vimin_s32_relu(dd[4], 255) | (__vimin_s32_relu(dd[5], 255) << 8) | (__vimin_s32_relu(dd[6], 255) << 16) |
            (__vimin_s32_relu(dd[7], 255) << 24);
      if (live) *reinterpret_cast<uint2*>(smem + pbase[c] + (by * 8 + r) * pitch[c] + bx * 8) = o;
    }
  }
  __syncthreads();

  // ---- upsample + colour + store
  const int H = d.height, W = d.width;
  const int y_lo = m0 * 8 * g.vmax, y_hi = min(m1 * 8 * g.vmax, H);
  uint8_t* img = dst + d.dst_offset;
  // plane views indexable with absolute sample rows (the offsets may be negative; every access lands inside the plane)
  const int oY = pbase[0] - br0[0] * 8 * pitch[0], oB = pbase[1] - br0[1] * 8 * pitch[1], oR = pbase[2] - br0[2] * 8 * pitch[2];
  const uint8_t* PY = smem + oY;
  const uint8_t* PB = smem + oB;
  const uint8_t* PR = smem + oR;
  const int wp = (W + 1) >> 1;   // pixel pairs per row
  if (g.mode == 3) {
    // jdsample.c h2v2_fancy_upsample: vertical 3:1 blend with the nearer / further chroma row, horizontal 3:1 with rounding
    // 8 / 7, edge replication at the real image border (chroma plane of ceil(H/2) x ceil(W/2) samples)
    const int hc = (H + 1) >> 1, wc = wp;
    const int cy_lo = y_lo >> 1, cy_hi = (y_hi + 1) >> 1;
    for (int qi = tid; qi < (cy_hi - cy_lo) * wc; qi += kPixThreads) {
      const int cy = cy_lo + qi / wc, cx = qi % wc;
      const int o_m = max(cy - 1, 0), o_p = min(cy + 1, hc - 1);
      const int cxm = max(cx - 1, 0), cxp = min(cx + 1, wc - 1);
      int ch[2][2][2];
#pragma unroll
      for (int comp = 0; comp < 2; ++comp) {
        const uint8_t* P = comp ? PR : PB;
        const int pt = pitch[1];
        const int n_l = P[cy * pt + cxm], n_c = P[cy * pt + cx], n_r = P[cy * pt + cxp];
        const int u_l = 3 * n_l + P[o_m * pt + cxm], u_c = 3 * n_c + P[o_m * pt + cx], u_r = 3 * n_r + P[o_m * pt + cxp];
        const int d_l = 3 * n_l + P[o_p * pt + cxm], d_c = 3 * n_c + P[o_p * pt + cx], d_r = 3 * n_r + P[o_p * pt + cxp];
        ch[comp][0][0] = (3 * u_c + u_l + 8) >> 4;
        ch[comp][0][1] = (3 * u_c + u_r + 7) >> 4;
        ch[comp][1][0] = (3 * d_c + d_l + 8) >> 4;
        ch[comp][1][1] = (3 * d_c + d_r + 7) >> 4;
        // jdsample.c jinit_upsampler: planes of one or two samples per row are replicated (h2v2_upsample), not interpolated
        if (wc <= 2) ch[comp][0][0] = ch[comp][0][1] = ch[comp][1][0] = ch[comp][1][1] = n_c;
      }
      const int nx = (2 * cx + 1 < W) ? 2 : 1;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int y = 2 * cy + dy;
        if (y >= H) break;
        int px[6];
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
          ycc_to_rgb(PY[y * pitch[0] + 2 * cx + dx], ch[0][dy][dx], ch[1][dy][dx], px + 3 * dx);
        store_rgb_pair(img + ((size_t)y * W + 2 * cx) * 3, px, nx);
      }
    }
  } else {
    for (int qi = tid; qi < (y_hi - y_lo) * wp; qi += kPixThreads) {
      const int y = y_lo + qi / wp, cx = qi % wp;
      const int nx = (2 * cx + 1 < W) ? 2 : 1;
      int px[6];
      if (g.mode == 0) {
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) px[3 * dx] = px[3 * dx + 1] = px[3 * dx + 2] = PY[y * pitch[0] + 2 * cx + dx];
      } else if (g.mode == 1) {
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int x = 2 * cx + dx;
          ycc_to_rgb(PY[y * pitch[0] + x], PB[y * pitch[1] + x], PR[y * pitch[2] + x], px + 3 * dx);
        }
      } else {
        // jdsample.c h2v1_fancy_upsample: out[2i] = (3 c[i] + c[i-1] + 1) >> 2, out[2i+1] = (3 c[i] + c[i+1] + 2) >> 2, the
        // first and last output columns copy their sample (== the same formulas with a clamped neighbour)
        const int cxm = max(cx - 1, 0), cxp = min(cx + 1, wp - 1);
        const int b0 = PB[y * pitch[1] + cx], r0 = PR[y * pitch[2] + cx];
        const int cbl = (3 * b0 + PB[y * pitch[1] + cxm] + 1) >> 2, cbr = (3 * b0 + PB[y * pitch[1] + cxp] + 2) >> 2;
        const int crl = (3 * r0 + PR[y * pitch[2] + cxm] + 1) >> 2, crr = (3 * r0 + PR[y * pitch[2] + cxp] + 2) >> 2;
        const bool fancy = wp > 2;      // jinit_upsampler: h2v1_upsample (replication) for planes of one or two samples per row
        ycc_to_rgb(PY[y * pitch[0] + 2 * cx], fancy ? cbl : b0, fancy ? crl : r0, px);
        ycc_to_rgb(PY[y * pitch[0] + 2 * cx + 1], fancy ? cbr : b0, fancy ? crr : r0, px + 3);
      }
      store_rgb_pair(img + ((size_t)y * W + 2 * cx) * 3, px, nx);
    }
  }
}

// ---- host: header parser -------------------------------------------------------------------------------------------------
inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

}  // namespace
}  // namespace vip

extern "C" int vip_jpeg_parse(const uint8_t* f, size_t len, vip_jpeg_desc* d) {
  using namespace vip;
  VIP_REQUIRE(d != nullptr && (f != nullptr || len == 0), VIP_ERR_INVALID, "vip_jpeg_parse: null argument");
  memset(d, 0, sizeof(*d));
  d->status = VIP_JPEG_NOT_JPEG;
  if (len < 4 || f[0] != 0xFF || f[1] != 0xD8) return VIP_OK;
  size_t p = 2;
  bool have_frame = false, have_qt[4] = {false, false, false, false}, have_ht[4] = {false, false, false, false};
  bool unsupported = false;
  int comp_id[3] = {0, 0, 0};
  int adobe_transform = -1;
  while (p + 4 <= len) {
    if (f[p] != 0xFF) return VIP_OK;                         // garbage between segments
    while (p < len && f[p] == 0xFF) ++p;                      // fill bytes
    if (p >= len) return VIP_OK;
    const int m = f[p++];
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (m == 0xD9) return VIP_OK;                             // EOI before any scan
    if (p + 2 > len) return VIP_OK;
    const size_t L = (size_t)be16(f + p);
    if (L < 2 || p + L > len) return VIP_OK;
    const uint8_t* s = f + p + 2;
    const size_t n = L - 2;
    if (m == 0xDB) {                                          // DQT (T.81 B.2.4.1)
      size_t i = 0;
      while (i < n) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        ++i;
        if (tq > 3 || i + (pq ? 128 : 64) > n) return VIP_OK;
        for (int k = 0; k < 64; ++k) {
          const int v = pq ? be16(s + i + 2 * k) : s[i + k];
          d->qt[tq][h_zigzag[k]] = (uint16_t)v;
        }
        if (pq) unsupported = true;                            // 16-bit tables: 12-bit data
        have_qt[tq] = true;
        i += pq ? 128 : 64;
      }
    } else if (m == 0xC4) {                                   // DHT (B.2.4.2)
      size_t i = 0;
      while (i < n) {
        if (i + 17 > n) return VIP_OK;
        const int tc = s[i] >> 4, th = s[i] & 15;
        int total = 0;
        for (int k = 0; k < 16; ++k) total += s[i + 1 + k];
        if (tc > 1 || total > 256 || i + 17 + total > n) return VIP_OK;
        if (th > 1) {
          unsupported = true;                                  // baseline allows two tables per class
        } else {
          const int t = tc * 2 + th;
          memcpy(d->huff_bits[t], s + i + 1, 16);
          memset(d->huff_vals[t], 0, 256);
          memcpy(d->huff_vals[t], s + i + 17, total);
          have_ht[t] = true;
        }
        i += 17 + total;
      }
    } else if (m == 0xC0 || m == 0xC1) {                      // SOF0 / SOF1: sequential Huffman (B.2.2)
      if (n < 6) return VIP_OK;
      const int prec = s[0];
      d->height = be16(s + 1);
      d->width = be16(s + 3);
      d->ncomp = s[5];
      if (prec != 8) unsupported = true;
      if (d->height <= 0 || d->width <= 0) return VIP_OK;
      if (d->ncomp != 1 && d->ncomp != 3) {
        unsupported = true;
      } else {
        if (n < (size_t)(6 + 3 * d->ncomp)) return VIP_OK;
        for (int c = 0; c < d->ncomp; ++c) {
          comp_id[c] = s[6 + 3 * c];
          d->hs[c] = s[7 + 3 * c] >> 4;
          d->vs[c] = s[7 + 3 * c] & 15;
          d->tq[c] = s[8 + 3 * c];
          if (d->tq[c] > 3) return VIP_OK;
        }
      }
      have_frame = true;
    } else if ((m >= 0xC2 && m <= 0xCF) && m != 0xC8) {       // progressive / lossless / arithmetic / hierarchical frames
      if (n >= 5) {
        d->height = be16(s + 1);
        d->width = be16(s + 3);
      }
      d->status = VIP_JPEG_UNSUPPORTED;
      return VIP_OK;
    } else if (m == 0xDD) {                                   // DRI
      if (n < 2) return VIP_OK;
      d->restart_interval = be16(s);
    } else if (m == 0xEE) {                                   // APP14 "Adobe": colour transform flag
      if (n >= 12 && memcmp(s, "Adobe", 5) == 0) adobe_transform = s[11];
    } else if (m == 0xDA) {                                   // SOS (B.2.3)
      if (!have_frame) return VIP_OK;
      d->status = VIP_JPEG_UNSUPPORTED;
      if (unsupported) return VIP_OK;
      if (n < 1) return VIP_OK;
      const int ns = s[0];
      if (ns != d->ncomp || n < (size_t)(4 + 2 * ns)) return VIP_OK;          // multi-scan (non-interleaved) files
      for (int c = 0; c < ns; ++c) {
        if (s[1 + 2 * c] != comp_id[c]) return VIP_OK;
        d->td[c] = s[2 + 2 * c] >> 4;
        d->ta[c] = s[2 + 2 * c] & 15;
        if (d->td[c] > 1 || d->ta[c] > 1 || !have_ht[d->td[c]] || !have_ht[2 + d->ta[c]] || !have_qt[d->tq[c]]) return VIP_OK;
      }
      if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63 || s[3 + 2 * ns] != 0) return VIP_OK;
      if (d->ncomp == 1) {
        d->hs[0] = d->vs[0] = 1;                              // A.2.2: a single-component scan is not interleaved
      } else {
        // libjpeg's colour space guess (jdapimin.c default_decompress_parms): Adobe transform 0 or ids 'R','G','B' = RGB
        if (adobe_transform == 0 || (adobe_transform < 0 && comp_id[0] == 'R' && comp_id[1] == 'G' && comp_id[2] == 'B'))
          return VIP_OK;
        if (d->hs[1] != 1 || d->vs[1] != 1 || d->hs[2] != 1 || d->vs[2] != 1) return VIP_OK;
        const int h = d->hs[0], v = d->vs[0];
        if (!((h == 1 && v == 1) || (h == 2 && v == 1) || (h == 2 && v == 2))) return VIP_OK;
      }
      if ((long long)d->width > 8192 || (long long)d->height > 8192) return VIP_OK;
      // entropy-coded segment: up to the first marker that is neither a stuffed zero nor RSTn
      const size_t start = p + L;
      size_t q = start;
      while (q < len) {
        const uint8_t* ff = static_cast<const uint8_t*>(memchr(f + q, 0xFF, len - q));
        if (ff == nullptr) { q = len; break; }
        q = (size_t)(ff - f);
        if (q + 1 >= len) { q = len; break; }
        const int nx = f[q + 1];
        if (nx == 0x00 || (nx >= 0xD0 && nx <= 0xD7)) { q += 2; continue; }
        if (nx == 0xFF) { q += 1; continue; }
        break;
      }
      d->scan_offset = (int32_t)start;
      d->scan_bytes = (int32_t)(q - start);
      // a second scan after this one (q points at a marker other than EOI) would be a multi-scan file
      if (q + 1 < len && f[q + 1] != 0xD9) {
        // tolerate trailing non-scan segments?  libjpeg would keep decoding scans: be strict
        if (f[q + 1] == 0xDA || f[q + 1] == 0xC4 || f[q + 1] == 0xDB) return VIP_OK;
      }
      d->status = VIP_JPEG_OK;
      return VIP_OK;
    }
    p += L;
  }
  return VIP_OK;
}

extern "C" int vip_jpeg_plan(vip_jpeg_desc* descs, int N, int64_t* dst_bytes, int64_t* coef_blocks) {
  using namespace vip;
  VIP_REQUIRE(N >= 0 && (descs != nullptr || N == 0), VIP_ERR_INVALID, "vip_jpeg_plan: bad arguments");
  int64_t db = 0, cb = 0;
  for (int i = 0; i < N; ++i) {
    vip_jpeg_desc& d = descs[i];
    d.dst_offset = db;
    d.coef_offset = cb;
    if (d.width > 0 && d.height > 0) db += (int64_t)d.width * d.height * 3;
    if (d.status == VIP_JPEG_OK) cb += geometry(d).nblocks;
  }
  if (dst_bytes != nullptr) *dst_bytes = db;
  if (coef_blocks != nullptr) *coef_blocks = cb;
  return VIP_OK;
}

extern "C" int vip_jpeg_decode(const uint8_t* data, const vip_jpeg_desc* descs_host, const vip_jpeg_desc* descs_dev, int N,
                               int16_t* coef, uint8_t* dst, int32_t* err, void* cuda_stream) {
  using namespace vip;
  VIP_REQUIRE(N >= 0, VIP_ERR_INVALID, "vip_jpeg_decode: N < 0");
  if (N == 0) return VIP_OK;
  VIP_REQUIRE(data != nullptr && descs_host != nullptr && descs_dev != nullptr && dst != nullptr, VIP_ERR_INVALID,
              "vip_jpeg_decode: null argument");
  int bands = 0, smem = 0, live = 0;
  for (int i = 0; i < N; ++i) {
    if (descs_host[i].status != VIP_JPEG_OK) continue;
    ++live;
    const Geo g = geometry(descs_host[i]);
    const int R = band_rows(g);
    bands = std::max(bands, (g.mcuy + R - 1) / R);
    smem = std::max(smem, R * g.band_bytes_per_row + g.band_halo_bytes);
  }
  if (live == 0) return VIP_OK;
  VIP_REQUIRE(coef != nullptr, VIP_ERR_INVALID, "vip_jpeg_decode: null coefficient workspace");
  VIP_REQUIRE((reinterpret_cast<uintptr_t>(coef) & 15) == 0, VIP_ERR_INVALID, "vip_jpeg_decode: coef must be 16-byte aligned");
  VIP_REQUIRE(smem <= 200 * 1024, VIP_ERR_UNSUPPORTED, "vip_jpeg_decode: image too wide (%d bytes of planes per MCU row)", smem);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  VIP_LAUNCH((jpeg_entropy_kernel), N, 32, 0, st, data, descs_dev, coef, err);
  VIP_CUDA(cudaGetLastError());
  VIP_CUDA(cudaFuncSetAttribute(jpeg_pixels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  VIP_LAUNCH((jpeg_pixels_kernel), dim3(bands, N), kPixThreads, smem, st, descs_dev, coef, dst);
  VIP_CUDA(cudaGetLastError());
  count_launch(2);
  return VIP_OK;
}
