// Baseline-JPEG decode on the GPU, bit-exact with the reference's loader.
//
// The reference decodes every frame / flow image with Pillow's Image.open (Sheet03/spatialModel.py:76-79,
// temporalModel.py:85-88) on files written by cv2.imwrite (utils.py:116-120): 8-bit baseline sequential JPEG, YCbCr 4:2:0
// (frames) or one component (flow images), Huffman coded.  Pillow's libjpeg-turbo default path is reproduced step by
// step so that the pixels entering K1 are the reference's pixels:
//   jpeg_huffman_kernel   jdhuff.c decode_mcu -- canonical codes (maxcode / valoffset), RECEIVE/EXTEND, DC prediction,
//                         ZRL / EOB, 0xFF00 unstuffing, restart markers.  The bit stream of an image is inherently serial:
//                         one THREAD per image (a batch has 10^3..10^4 images; 1..32 images per warp depending on the
//                         batch size), non-zero coefficients scattered into a zero-filled int16 workspace.
//   jpeg_idct_kernel      dequantise + jidctint.c jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2) + range limit: one thread
//                         per 8x8 block.  One-component images are written straight into the image store.
//   jpeg_color_kernel     jdsample.c h2v2_fancy_upsample (triangle filter, biases 8/7, replicated context rows; plain
//                         replication when the chroma plane is <= 2 samples wide) + jdcolor.c ycc_rgb_convert
//                         (16-bit fixed point): one thread per output pixel, interleaved RGB out.
// Header parsing (markers, DQT/DHT/SOF0/SOS/DRI) is host work (video_analytics_b200/jpeg.py).
#include "va_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace va {

struct JpegImage {            // mirrors va_jpeg_image (include/va_b200.h) field for field
  unsigned long long scan_offset;    // first entropy-coded byte, relative to `bitstreams`
  unsigned long long out_offset;     // byte offset of the decoded image [H][W][n_comp] in `out`
  unsigned int scan_bytes;           // bytes from scan_offset to the end of the file
  unsigned int restart_interval;     // MCUs between RSTn markers, 0 = none
  unsigned short width, height;
  unsigned char n_comp;              // 1 or 3
  unsigned char sampling;            // 0 = one component, 1 = 4:4:4, 2 = 4:2:0
  unsigned char qt[3], dc[3], ac[3]; // table indices per component
  unsigned char pad[9];
};
static_assert(sizeof(JpegImage) == 48, "JpegImage layout");

struct JpegHuff {             // derived table (jpeg_make_d_derived_tbl)
  int maxcode[18];            // largest code of length l (-1 if none); [17] = sentinel
  int valoffset[17];          // huffval index = code + valoffset[l]
  unsigned char huffval[256];
  int pad;
};
static_assert(sizeof(JpegHuff) == 400, "JpegHuff layout");

struct JpegGeom {             // per image, computed on the host side of va_jpeg_decode
  unsigned long long coef_block0;    // first block of this image in the coefficient workspace
  unsigned long long plane_offset;   // byte offset of this image's component planes (colour images only)
  unsigned long long ustream_offset; // byte offset (multiple of 4) of this image's unstuffed scan data (parallel decoder)
  int mcux, mcuy;                    // MCUs per row / column
  int bw[3], bh[3];                  // blocks per row / column of each component plane
};

__constant__ unsigned char c_natural[80] = {
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
    62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};      // + 16 safety entries (libjpeg)

// ------------------------------------------------------------------------------------------------ entropy decoding
// One thread decodes one image; a warp therefore runs 32 independent serial decoders, and what matters is (a) the
// dependent-instruction latency per symbol and (b) that the lanes stay converged.  Both are served by a BRANCH-FREE code
// length search: canonical codes of length l occupy the 16-bit-left-aligned interval below bound[l] = (maxcode[l]+1) <<
// (16-l) (unused lengths inherit the previous bound), so  l = 1 + #{ l' : look >= bound[l'] }  -- 16 independent
// shared-memory reads and compares instead of a data-dependent loop over global memory.
constexpr int kJpegMaxTables = 8;            // Huffman tables kept in shared memory per block

struct HuffSmem {
  unsigned int bound[kJpegMaxTables][17];    // [t][l], l = 1..16; [0] unused
  int valoffset[kJpegMaxTables][17];
  unsigned char huffval[kJpegMaxTables][256];
  unsigned char natural[80];
};

struct BitReader {
  const unsigned char* p;
  const unsigned char* end;
  unsigned long long buf;
  int nbits;
  bool marker;

  __device__ __forceinline__ void fill_slow() {            // byte-wise: 0xFF00 unstuffing, stop at markers
    while (nbits <= 56) {
      unsigned int c = 0;
      if (!marker && p < end) {
        c = __ldg(p);
        if (c == 0xFF) {
          const unsigned int c2 = (p + 1 < end) ? __ldg(p + 1) : 0xD9u;
          if (c2 == 0) p += 2;
          else { c = 0; marker = true; }          // a marker: feed zero bits from here on (jpeg_fill_bit_buffer)
        } else {
          ++p;
        }
      }
      buf = (buf << 8) | c;
      nbits += 8;
    }
  }
  // keep at least 32 valid bits: four independent byte loads, inserted at once unless one of them is 0xFF
  __device__ __forceinline__ void refill() {
    if (nbits >= 32) return;
    if (!marker && p + 4 <= end) {
      const unsigned int b0 = __ldg(p), b1 = __ldg(p + 1), b2 = __ldg(p + 2), b3 = __ldg(p + 3);
      if (b0 != 0xFF && b1 != 0xFF && b2 != 0xFF && b3 != 0xFF) {
        buf = (buf << 32) | (unsigned long long)((b0 << 24) | (b1 << 16) | (b2 << 8) | b3);
        nbits += 32;
        p += 4;
        return;
      }
    }
    fill_slow();
  }
  __device__ __forceinline__ unsigned int get(int n) {     // n <= 16; caller keeps nbits >= 32 via refill()
    nbits -= n;
    return (unsigned int)(buf >> nbits) & ((1u << n) - 1u);
  }
  __device__ __forceinline__ int decode(const HuffSmem& h, int t) {
    refill();
    const unsigned int look = (unsigned int)(buf >> (nbits - 16)) & 0xFFFFu;
    int l = 1;
#pragma unroll
    for (int k = 1; k <= 16; ++k) l += (look >= h.bound[t][k]) ? 1 : 0;
    if (l > 16) { nbits -= 16; return 0; }        // corrupt data: libjpeg warns and returns 0
    nbits -= l;
    const int code = (int)(look >> (16 - l));
    return (int)h.huffval[t][(code + h.valoffset[t][l]) & 255];
  }
  __device__ __forceinline__ void restart() {     // process_restart: drop the partial byte, skip RSTn
    nbits = 0; buf = 0; marker = false;
    while (p + 1 < end && !(__ldg(p) == 0xFF && __ldg(p + 1) >= 0xD0 && __ldg(p + 1) <= 0xD7)) ++p;
    p += 2;
  }
};

__device__ __forceinline__ int jpeg_extend(unsigned int v, int s) {
  return (int)v < (1 << (s - 1)) ? (int)v - (1 << s) + 1 : (int)v;
}

__device__ __forceinline__ void huffman_decode_images(const unsigned char* __restrict__ bitstreams,
                                                      const JpegImage* __restrict__ images,
                                                      const JpegGeom* __restrict__ geom, const HuffSmem& h, int n_images,
                                                      int lane_stride, short* __restrict__ coef) {
  // lane_stride = 32 / images per warp: with fewer images per warp the decoders of a warp diverge less (1 = none)
  if (threadIdx.x % lane_stride) return;
  const int img = blockIdx.x * (blockDim.x / lane_stride) + threadIdx.x / lane_stride;
  if (img >= n_images) return;
  const JpegImage im = images[img];
  const JpegGeom g = geom[img];
  BitReader br;
  br.p = bitstreams + im.scan_offset;
  br.end = br.p + im.scan_bytes;
  br.buf = 0; br.nbits = 0; br.marker = false;
  int pred[3] = {0, 0, 0};
  // blocks of component c inside one MCU: (hs x vs)
  const int hs0 = im.sampling == 2 ? 2 : 1, vs0 = hs0;
  unsigned long long plane0[3];
  plane0[0] = g.coef_block0;
  plane0[1] = plane0[0] + (unsigned long long)g.bw[0] * g.bh[0];
  plane0[2] = plane0[1] + (unsigned long long)g.bw[1] * g.bh[1];
  unsigned int todo = im.restart_interval;
  for (int my = 0; my < g.mcuy; ++my) {
    for (int mx = 0; mx < g.mcux; ++mx) {
      if (im.restart_interval) {
        if (todo == 0) {
          br.restart();
          pred[0] = pred[1] = pred[2] = 0;
          todo = im.restart_interval;
        }
        --todo;
      }
      for (int c = 0; c < im.n_comp; ++c) {
        const int hs = c == 0 ? hs0 : 1, vs = c == 0 ? vs0 : 1;
        const int dct = im.dc[c], act = im.ac[c];
        for (int by = 0; by < vs; ++by) {
          for (int bx = 0; bx < hs; ++bx) {
            short* blk = coef + (plane0[c] + (unsigned long long)(my * vs + by) * g.bw[c] + (mx * hs + bx)) * 64;
            int s = br.decode(h, dct);
            if (s) { br.refill(); pred[c] += jpeg_extend(br.get(s), s); }
            blk[0] = (short)pred[c];
            int k = 1;
            while (k < 64) {
              const int rs = br.decode(h, act);
              const int r = rs >> 4;
              s = rs & 15;
              if (s) {
                k += r;
                br.refill();
                blk[h.natural[k < 80 ? k : 79]] = (short)jpeg_extend(br.get(s), s);
                ++k;
              } else if (r == 15) {
                k += 16;
              } else {
                break;
              }
            }
          }
        }
      }
    }
  }
}

// Decoding tables in the layout the decoder reads (bound / valoffset / symbols / zigzag), built once per call in global
// memory.  USE_SMEM = false (VA_JPEG_SMEM_TABLES=0) reads them through L1 instead of copying them to shared memory: 15-20 %
// slower alone.  It was written to let decoder blocks co-reside with the persistent layer kernels on a side stream, but
// those take 227 KB + the 1 KB system reservation of the SM's 228 KB, so NO other block fits next to them whatever it
// needs: a decode issued ahead on a side stream runs in the gaps between layer launches (measured: bench e2e_jpeg).
__global__ void jpeg_tables_kernel(const JpegHuff* __restrict__ htab, int n_tables, HuffSmem* __restrict__ out) {
  for (int idx = threadIdx.x; idx < n_tables * 17; idx += blockDim.x) out->valoffset[idx / 17][idx % 17] = htab[idx / 17].valoffset[idx % 17];
  for (int idx = threadIdx.x; idx < n_tables * 256; idx += blockDim.x) out->huffval[idx >> 8][idx & 255] = htab[idx >> 8].huffval[idx & 255];
  for (int idx = threadIdx.x; idx < 80; idx += blockDim.x) out->natural[idx] = c_natural[idx];
  if (threadIdx.x < n_tables) {
    const int t = threadIdx.x;
    unsigned int prev = 0;
    out->bound[t][0] = 0;
    for (int l = 1; l <= 16; ++l) {
      const int mc = htab[t].maxcode[l];
      if (mc >= 0) prev = (unsigned int)(mc + 1) << (16 - l);
      out->bound[t][l] = prev;
    }
  }
}

template <bool USE_SMEM>
__global__ void __launch_bounds__(32) jpeg_huffman_kernel(const unsigned char* __restrict__ bitstreams,
                                                          const JpegImage* __restrict__ images,
                                                          const JpegGeom* __restrict__ geom,
                                                          const JpegHuff* __restrict__ htab,
                                                          const HuffSmem* __restrict__ gtab, int n_tables, int n_images,
                                                          int lane_stride, short* __restrict__ coef) {
  if (!USE_SMEM) {
    huffman_decode_images(bitstreams, images, geom, *gtab, n_images, lane_stride, coef);
    return;
  }
  __shared__ HuffSmem h;
  for (int idx = threadIdx.x; idx < n_tables * 17; idx += blockDim.x) {
    const int t = idx / 17, l = idx % 17;
    h.valoffset[t][l] = htab[t].valoffset[l];
  }
  for (int idx = threadIdx.x; idx < n_tables * 256; idx += blockDim.x) h.huffval[idx >> 8][idx & 255] = htab[idx >> 8].huffval[idx & 255];
  for (int idx = threadIdx.x; idx < 80; idx += blockDim.x) h.natural[idx] = c_natural[idx];
  if (threadIdx.x < n_tables) {
    const int t = threadIdx.x;
    unsigned int prev = 0;
    h.bound[t][0] = 0;
    for (int l = 1; l <= 16; ++l) {
      const int mc = htab[t].maxcode[l];
      if (mc >= 0) prev = (unsigned int)(mc + 1) << (16 - l);
      h.bound[t][l] = prev;
    }
  }
  __syncthreads();
  huffman_decode_images(bitstreams, images, geom, h, n_images, lane_stride, coef);
}

// ------------------------------------------------------------------------------------------------ parallel entropy decoding
// One thread per image leaves a call at the latency of one serial decode (~20 ms for a 67 KB file) however few images
// it has.  For small batches an image is decoded by a whole thread block instead, exploiting that Huffman streams
// self-synchronise (Klein & Wiseman; Weissenberger & Schmidt for JPEG on GPUs):
//   0. the block removes the 0xFF00 byte stuffing (block-wide compaction) so that bit positions are addressable;
//   1. the stream is cut into chunks of S bits; every thread decodes its chunks from the chunk start ASSUMING a block
//      boundary -- wrong for most chunks, but after a few symbols the decoder falls into step with the true symbol and
//      block boundaries -- and records where it stopped (first symbol start at/after the chunk end), in which decoder
//      state (coefficient index k, block-in-MCU b) and how many blocks it completed;
//   2. every chunk is decoded again from its PREDECESSOR's recorded end state; repeated until no record changes.  Chunk 0
//      starts from the true state, so the fixed point is the true segmentation (normally reached after 1-2 rounds);
//   3. an exclusive scan of the per-chunk block counts gives every chunk its first block index; a last decoding pass
//      writes the coefficients (DC as differences);
//   4. a scan per component over the blocks in stream order turns the DC differences into DC values.
// ~4 passes of work per image instead of 1, but spread over 256 threads: 206 images in ~1.5 ms instead of 22 ms.
constexpr int kParThreads = 256;
constexpr int kParMaxChunks = 512;      // 8 KB of records: small enough to co-reside with a layer kernel that leaves ~17 KB

struct ChunkRec {
  unsigned int p;          // bit position of the first symbol that starts at or after the chunk end
  unsigned short state;    // k | b << 8
  unsigned short nblocks;  // blocks completed by symbols that start inside this chunk
};

__device__ __forceinline__ unsigned int peek32(const unsigned int* __restrict__ w, unsigned int p) {
  const unsigned int i = p >> 5;
  const unsigned int hi = __byte_perm(w[i], 0, 0x0123), lo = __byte_perm(w[i + 1], 0, 0x0123);
  return __funnelshift_l(lo, hi, p & 31);
}

// Decode the symbols that start in [p, limit) of an image whose decoder state at p is (k, b).  WRITE: emit coefficients
// for block indices first_block.. (DC as a difference), never beyond total_blocks.
template <bool WRITE>
__device__ __forceinline__ ChunkRec decode_chunk(const unsigned int* __restrict__ w, unsigned int p, unsigned int limit, int k, int b,
                                                 const HuffSmem& h, const JpegImage& im, int blocks_per_mcu, const JpegGeom& g,
                                                 unsigned int first_block, unsigned int total_blocks, short* __restrict__ coef) {
  unsigned int nb = 0;
  short* blk = nullptr;
  unsigned int cur = first_block;
  auto block_ptr = [&](unsigned int sb) -> short* {
    // stream block index -> coefficient workspace position (MCU interleaving of the scan)
    unsigned long long pos;
    if (blocks_per_mcu == 1) {
      pos = g.coef_block0 + sb;
    } else {
      const unsigned int mcu = sb / blocks_per_mcu, j = sb - mcu * blocks_per_mcu;
      const unsigned int my = mcu / g.mcux, mx = mcu - my * g.mcux;
      const unsigned long long ny = (unsigned long long)g.bw[0] * g.bh[0], nc = (unsigned long long)g.bw[1] * g.bh[1];
      if (blocks_per_mcu == 6) {
        if (j < 4) pos = g.coef_block0 + (unsigned long long)(my * 2 + (j >> 1)) * g.bw[0] + (mx * 2 + (j & 1));
        else pos = g.coef_block0 + ny + (j - 4) * nc + (unsigned long long)my * g.bw[1] + mx;
      } else {
        pos = g.coef_block0 + (j == 0 ? 0ull : ny + (j - 1) * nc) + (unsigned long long)my * g.bw[j ? 1 : 0] + mx;
      }
    }
    return coef + pos * 64;
  };
  if (WRITE && cur < total_blocks) blk = block_ptr(cur);
  while (p < limit) {
    const int comp = blocks_per_mcu == 1 ? 0 : (blocks_per_mcu == 6 ? (b < 4 ? 0 : b - 3) : b);
    const int t = k == 0 ? im.dc[comp] : im.ac[comp];
    const unsigned int win = peek32(w, p);
    const unsigned int look = win >> 16;
    int l = 1;
#pragma unroll
    for (int q = 1; q <= 16; ++q) l += (look >= h.bound[t][q]) ? 1 : 0;
    int sym = 0;
    if (l > 16) l = 16;
    else sym = h.huffval[t][((int)(look >> (16 - l)) + h.valoffset[t][l]) & 255];
    if (k == 0) {                                   // DC difference
      const int sbits = sym & 15;
      if (WRITE && blk) blk[0] = sbits ? (short)jpeg_extend((win << l) >> (32 - sbits), sbits) : (short)0;
      p += l + sbits;
      k = 1;
    } else {
      const int r = sym >> 4, sbits = sym & 15;
      if (sbits) {
        k += r;
        if (WRITE && blk) blk[h.natural[k < 80 ? k : 79]] = (short)jpeg_extend((win << l) >> (32 - sbits), sbits);
        ++k;
        p += l + sbits;
      } else {
        k = r == 15 ? k + 16 : 64;                  // ZRL / EOB
        p += l;
      }
    }
    if (k >= 64) {                                  // block complete
      k = 0;
      if (++b == blocks_per_mcu) b = 0;
      ++nb;
      if (WRITE) {
        ++cur;
        blk = cur < total_blocks ? block_ptr(cur) : nullptr;
      }
    }
  }
  ChunkRec rec;
  rec.p = p;
  rec.state = (unsigned short)(k | (b << 8));
  rec.nblocks = (unsigned short)(nb > 0xFFFF ? 0xFFFF : nb);
  return rec;
}

__global__ void __launch_bounds__(kParThreads) jpeg_huffman_parallel_kernel(const unsigned char* __restrict__ bitstreams,
                                                                            const JpegImage* __restrict__ images,
                                                                            const JpegGeom* __restrict__ geom,
                                                                            const HuffSmem* __restrict__ gtab,
                                                                            unsigned char* __restrict__ ustream,
                                                                            short* __restrict__ coef) {
  __shared__ HuffSmem h;
  __shared__ ChunkRec rec[2][kParMaxChunks];
  __shared__ unsigned int s_scan[kParThreads];
  __shared__ unsigned int s_misc[4];
  const int tid = threadIdx.x;
  {
    const unsigned int* src = reinterpret_cast<const unsigned int*>(gtab);
    unsigned int* dst = reinterpret_cast<unsigned int*>(&h);
    for (int i = tid; i < (int)(sizeof(HuffSmem) / 4); i += kParThreads) dst[i] = src[i];
  }
  const JpegImage im = images[blockIdx.x];
  const JpegGeom g = geom[blockIdx.x];
  const unsigned char* in = bitstreams + im.scan_offset;
  unsigned char* ub = ustream + g.ustream_offset;        // 4-byte aligned (word reader), scan_bytes + 32 bytes long

  // ---- phase 0: remove byte stuffing, stop at the first marker
  if (tid == 0) { s_misc[0] = 0; s_misc[1] = 0xFFFFFFFFu; }
  __syncthreads();
  unsigned int out_base = 0;
  for (unsigned int base = 0; base < im.scan_bytes; base += kParThreads) {
    const unsigned int i = base + tid;
    unsigned int b = 0, prev = 0;
    bool valid = i < im.scan_bytes;
    if (valid) { b = in[i]; prev = i ? in[i - 1] : 0; }
    const bool is_marker2 = valid && prev == 0xFF && b != 0;                 // second byte of a marker: data ended at i-1
    if (is_marker2) atomicMin(&s_misc[1], i - 1);
    __syncthreads();
    const unsigned int end = s_misc[1];
    const bool keep = valid && i < end && !(b == 0 && prev == 0xFF);
    // block-wide exclusive scan of keep
    const unsigned int lane = tid & 31, wid = tid >> 5;
    const unsigned int ball = __ballot_sync(0xffffffffu, keep);
    const unsigned int within = __popc(ball & ((1u << lane) - 1u));
    if (lane == 0) s_scan[wid] = __popc(ball);
    __syncthreads();
    unsigned int woff = 0, total = 0;
    for (int q = 0; q < kParThreads / 32; ++q) { const unsigned int c = s_scan[q]; if (q < (int)wid) woff += c; total += c; }
    if (keep) ub[out_base + woff + within] = (unsigned char)b;
    out_base += total;
    __syncthreads();
    if (end != 0xFFFFFFFFu) break;
  }
  for (int i = tid; i < 16; i += kParThreads) ub[out_base + i] = 0;           // zero padding for the word reader
  __syncthreads();
  const unsigned int nbits = out_base * 8;
  const unsigned int* w = reinterpret_cast<const unsigned int*>(ub);
  const int blocks_per_mcu = im.n_comp == 1 ? 1 : (im.sampling == 2 ? 6 : 3);
  const unsigned int total_blocks = (unsigned int)((unsigned long long)g.bw[0] * g.bh[0] + 2ull * g.bw[1] * g.bh[1]);
  // chunk size: at least 1024 bits, at most kParMaxChunks chunks
  unsigned int S = 1024;
  while ((nbits + S - 1) / S > kParMaxChunks) S *= 2;
  const int nchunks = (int)((nbits + S - 1) / S);
  if (nchunks == 0) return;

  // ---- phase 1: every chunk from its own start, assuming a block boundary
  for (int c = tid; c < nchunks; c += kParThreads) {
    const unsigned int lim = min((unsigned int)(c + 1) * S, nbits);
    rec[0][c] = decode_chunk<false>(w, (unsigned int)c * S, lim, 0, 0, h, im, blocks_per_mcu, g, 0, 0, nullptr);
  }
  __syncthreads();
  // ---- phase 2: from the predecessor's end state until nothing changes
  int cur = 0;
  for (int round = 0; round < nchunks; ++round) {
    int changed = 0;
    for (int c = tid; c < nchunks; c += kParThreads) {
      ChunkRec r;
      if (c == 0) {
        r = rec[cur][0];
      } else {
        const ChunkRec pr = rec[cur][c - 1];
        const unsigned int lim = min((unsigned int)(c + 1) * S, nbits);
        r = decode_chunk<false>(w, pr.p, lim, pr.state & 255, pr.state >> 8, h, im, blocks_per_mcu, g, 0, 0, nullptr);
        const ChunkRec old = rec[cur][c];
        changed |= (r.p != old.p) | (r.state != old.state) | (r.nblocks != old.nblocks);
      }
      rec[cur ^ 1][c] = r;
    }
    cur ^= 1;
    if (!__syncthreads_or(changed)) break;
  }
  // ---- phase 3: exclusive scan of the block counts, then the writing pass
  unsigned int carry = 0;
  for (int base = 0; base < nchunks; base += kParThreads) {
    const int c = base + tid;
    const unsigned int v = c < nchunks ? rec[cur][c].nblocks : 0;
    // block-wide inclusive scan
    unsigned int x = v;
    const unsigned int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (unsigned)o) x += y; }
    if (lane == 31) s_scan[wid] = x;
    __syncthreads();
    unsigned int woff = 0, total = 0;
    for (int q = 0; q < kParThreads / 32; ++q) { const unsigned int cc = s_scan[q]; if (q < (int)wid) woff += cc; total += cc; }
    const unsigned int excl = carry + woff + x - v;
    if (c < nchunks) {
      const unsigned int p0 = c == 0 ? 0u : rec[cur][c - 1].p;
      const int st = c == 0 ? 0 : rec[cur][c - 1].state;
      const unsigned int lim = min((unsigned int)(c + 1) * S, nbits);
      decode_chunk<true>(w, p0, lim, st & 255, st >> 8, h, im, blocks_per_mcu, g, excl, total_blocks, coef);
    }
    carry += total;
    __syncthreads();
  }
  __syncthreads();
  // ---- phase 4: DC differences -> DC values, one running sum per component over the blocks in stream order
  for (int comp = 0; comp < im.n_comp; ++comp) {
    const unsigned int per_mcu = (blocks_per_mcu == 6 && comp == 0) ? 4 : 1;
    const unsigned int first_j = blocks_per_mcu == 1 ? 0 : (blocks_per_mcu == 6 ? (comp == 0 ? 0 : 3 + comp) : comp);
    const unsigned int ncomp_blocks = (unsigned int)((unsigned long long)g.bw[comp ? 1 : 0] * g.bh[comp ? 1 : 0]);
    int run = 0;
    for (unsigned int base = 0; base < ncomp_blocks; base += kParThreads) {
      const unsigned int n = base + tid;
      short* bp = nullptr;
      int v = 0;
      if (n < ncomp_blocks) {
        const unsigned int sb = blocks_per_mcu == 1 ? n : (n / per_mcu) * blocks_per_mcu + first_j + (n % per_mcu);
        // same mapping as decode_chunk's block_ptr
        unsigned long long pos;
        if (blocks_per_mcu == 1) {
          pos = g.coef_block0 + sb;
        } else {
          const unsigned int mcu = sb / blocks_per_mcu, j = sb - mcu * blocks_per_mcu;
          const unsigned int my = mcu / g.mcux, mx = mcu - my * g.mcux;
          const unsigned long long ny = (unsigned long long)g.bw[0] * g.bh[0], nc = (unsigned long long)g.bw[1] * g.bh[1];
          if (blocks_per_mcu == 6) {
            if (j < 4) pos = g.coef_block0 + (unsigned long long)(my * 2 + (j >> 1)) * g.bw[0] + (mx * 2 + (j & 1));
            else pos = g.coef_block0 + ny + (j - 4) * nc + (unsigned long long)my * g.bw[1] + mx;
          } else {
            pos = g.coef_block0 + (j == 0 ? 0ull : ny + (j - 1) * nc) + (unsigned long long)my * g.bw[j ? 1 : 0] + mx;
          }
        }
        bp = coef + pos * 64;
        v = bp[0];
      }
      int x = v;
      const unsigned int lane = tid & 31, wid = tid >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (unsigned)o) x += y; }
      if (lane == 31) s_scan[wid] = (unsigned int)x;
      __syncthreads();
      int woff = 0, total = 0;
      for (int q = 0; q < kParThreads / 32; ++q) { const int cc = (int)s_scan[q]; if (q < (int)wid) woff += cc; total += cc; }
      if (bp) bp[0] = (short)(run + woff + x);
      run += total;
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------ IDCT
#define JF_0_298631336 2446
#define JF_0_390180644 3196
#define JF_0_541196100 4433
#define JF_0_765366865 6270
#define JF_0_899976223 7373
#define JF_1_175875602 9633
#define JF_1_501321110 12299
#define JF_1_847759065 15137
#define JF_1_961570560 16069
#define JF_2_053119869 16819
#define JF_2_562915447 20995
#define JF_3_072711026 25172

// one 1-D pass of jpeg_idct_islow on d[0..7] (stride 1 in registers); SHIFT = descale amount
template <int SHIFT>
__device__ __forceinline__ void idct8(int (&d)[8]) {
  int z2 = d[2], z3 = d[6];
  int z1 = (z2 + z3) * JF_0_541196100;
  const int tmp2 = z1 + z3 * (-JF_1_847759065);
  const int tmp3 = z1 + z2 * JF_0_765366865;
  const int tmp0 = (d[0] + d[4]) << 13;
  const int tmp1 = (d[0] - d[4]) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  int t0 = d[7], t1 = d[5], t2 = d[3], t3 = d[1];
  z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2;
  int z4 = t1 + t3;
  const int z5 = (z3 + z4) * JF_1_175875602;
  t0 *= JF_0_298631336; t1 *= JF_2_053119869; t2 *= JF_3_072711026; t3 *= JF_1_501321110;
  z1 *= -JF_0_899976223; z2 *= -JF_2_562915447;
  z3 = z3 * (-JF_1_961570560) + z5;
  z4 = z4 * (-JF_0_390180644) + z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  constexpr int R = 1 << (SHIFT - 1);
  d[0] = (tmp10 + t3 + R) >> SHIFT; d[7] = (tmp10 - t3 + R) >> SHIFT;
  d[1] = (tmp11 + t2 + R) >> SHIFT; d[6] = (tmp11 - t2 + R) >> SHIFT;
  d[2] = (tmp12 + t1 + R) >> SHIFT; d[5] = (tmp12 - t1 + R) >> SHIFT;
  d[3] = (tmp13 + t0 + R) >> SHIFT; d[4] = (tmp13 - t0 + R) >> SHIFT;
}

__global__ void __launch_bounds__(128) jpeg_idct_kernel(const short* __restrict__ coef, const JpegImage* __restrict__ images,
                                                        const JpegGeom* __restrict__ geom,
                                                        const unsigned short* __restrict__ qtab, int n_images,
                                                        unsigned long long total_blocks, unsigned char* __restrict__ planes,
                                                        unsigned char* __restrict__ out) {
  const unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= total_blocks) return;
  // image of this block: last image whose first block <= b
  int lo = 0, hi = n_images - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (geom[mid].coef_block0 <= b) lo = mid; else hi = mid - 1;
  }
  const JpegImage im = images[lo];
  const JpegGeom g = geom[lo];
  unsigned long long local = b - g.coef_block0;
  int c = 0;
  while (c < 2 && local >= (unsigned long long)g.bw[c] * g.bh[c]) { local -= (unsigned long long)g.bw[c] * g.bh[c]; ++c; }
  const int by = (int)(local / g.bw[c]), bx = (int)(local % g.bw[c]);
  const unsigned short* q = qtab + im.qt[c] * 64;
  const uint4* src = reinterpret_cast<const uint4*>(coef + b * 64);
  int ws[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint4 v = __ldg(src + r);
    const short* s = reinterpret_cast<const short*>(&v);
#pragma unroll
    for (int k = 0; k < 8; ++k) ws[r][k] = (int)s[k] * (int)__ldg(q + r * 8 + k);     // DEQUANTIZE
  }
  // pass 1: columns
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int col[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) col[r] = ws[r][k];
    idct8<13 - 2>(col);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[r][k] = col[r];
  }
  // pass 2: rows, range limit, store
  unsigned char* dst;
  int stride, wlim, hlim;
  if (im.n_comp == 1) {
    dst = out + im.out_offset;
    stride = im.width; wlim = im.width; hlim = im.height;
  } else {
    unsigned long long off = g.plane_offset;
    for (int k = 0; k < c; ++k) off += (unsigned long long)g.bw[k] * g.bh[k] * 64;
    dst = planes + off;
    stride = g.bw[c] * 8; wlim = stride; hlim = g.bh[c] * 8;
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    idct8<13 + 2 + 3>(ws[r]);
    const int y = by * 8 + r;
    if (y < hlim) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int x = bx * 8 + k;
        int v = ((ws[r][k] + 512) & 1023) - 512;          // range_limit[x & RANGE_MASK] ...
        v = min(max(v + 128, 0), 255);                    // ... == clamp(x + CENTERJSAMPLE) on that window
        if (x < wlim) dst[(size_t)y * stride + x] = (unsigned char)v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ upsample + colour
__device__ __forceinline__ int fancy_h2v2(const unsigned char* __restrict__ p, int stride, int dw, int dh, int y, int x) {
  const int i = y >> 1, j = x >> 1;
  if (dw <= 2) return p[(size_t)i * stride + j];                            // h2v2_upsample (replication)
  const int far_i = (y & 1) ? min(i + 1, dh - 1) : max(i - 1, 0);
  const unsigned char* r0 = p + (size_t)i * stride;
  const unsigned char* r1 = p + (size_t)far_i * stride;
  const int cur = 3 * r0[j] + r1[j];
  if (x & 1) {
    const int jn = min(j + 1, dw - 1);
    return (cur * 3 + (3 * r0[jn] + r1[jn]) + 7) >> 4;
  }
  const int jl = max(j - 1, 0);
  return (cur * 3 + (3 * r0[jl] + r1[jl]) + 8) >> 4;
}

__global__ void __launch_bounds__(256) jpeg_color_kernel(const int* __restrict__ color_ids, const JpegImage* __restrict__ images,
                                                         const JpegGeom* __restrict__ geom,
                                                         const unsigned char* __restrict__ planes,
                                                         unsigned char* __restrict__ out) {
  const int img = color_ids[blockIdx.y];
  const JpegImage im = images[img];
  const JpegGeom g = geom[img];
  const int W = im.width, H = im.height;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= W * H) return;
  const int y = pix / W, x = pix - y * W;
  const unsigned char* py = planes + g.plane_offset;
  const unsigned char* pcb = py + (size_t)g.bw[0] * g.bh[0] * 64;
  const unsigned char* pcr = pcb + (size_t)g.bw[1] * g.bh[1] * 64;
  const int yy = py[(size_t)y * (g.bw[0] * 8) + x];
  int cb, cr;
  if (im.sampling == 2) {
    const int dw = (W + 1) >> 1, dh = (H + 1) >> 1;
    cb = fancy_h2v2(pcb, g.bw[1] * 8, dw, dh, y, x);
    cr = fancy_h2v2(pcr, g.bw[2] * 8, dw, dh, y, x);
  } else {
    cb = pcb[(size_t)y * (g.bw[1] * 8) + x];
    cr = pcr[(size_t)y * (g.bw[2] * 8) + x];
  }
  // jdcolor.c build_ycc_rgb_table: FIX(x) = (int)(x * 65536 + 0.5), ONE_HALF = 32768, arithmetic right shifts
  const int xb = cb - 128, xr = cr - 128;
  const int r = yy + ((91881 * xr + 32768) >> 16);
  const int gg = yy + ((-22554 * xb + 32768 + (-46802) * xr) >> 16);
  const int bl = yy + ((116130 * xb + 32768) >> 16);
  unsigned char* o = out + im.out_offset + ((size_t)y * W + x) * 3;
  o[0] = (unsigned char)min(max(r, 0), 255);
  o[1] = (unsigned char)min(max(gg, 0), 255);
  o[2] = (unsigned char)min(max(bl, 0), 255);
}

// ------------------------------------------------------------------------------------------------ host entry
namespace {
thread_local char g_jerr[256];

// Ring of pinned host buffers for the per-call descriptor upload; a slot is reused once its copy has completed.
struct PinnedSlot {
  unsigned char* host = nullptr;
  size_t bytes = 0;
  cudaEvent_t done = nullptr;
};
PinnedSlot* pinned_slot(size_t bytes) {
  constexpr int kSlots = 8;
  static thread_local PinnedSlot ring[kSlots];
  static thread_local int next = 0;
  PinnedSlot& s = ring[next];
  next = (next + 1) % kSlots;
  if (s.done == nullptr && cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (s.host != nullptr && cudaEventSynchronize(s.done) != cudaSuccess) return nullptr;      // normally long complete
  if (s.bytes < bytes) {
    if (s.host) cudaFreeHost(s.host);
    s.host = nullptr;
    s.bytes = 0;
    const size_t want = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
    if (cudaMallocHost(&s.host, want) != cudaSuccess) return nullptr;
    s.bytes = want;
  }
  return &s;
}
}

const char* jpeg_decode_run(const unsigned char* bitstreams, const void* images_host, int n_images, const unsigned short* qtables_host,
                            int n_q, const void* htables_host, int n_h, unsigned char* out, cudaStream_t st) {
  if (n_images <= 0) return nullptr;
  const JpegImage* im = static_cast<const JpegImage*>(images_host);
  std::vector<JpegGeom> geom((size_t)n_images);
  std::vector<int> color_ids;
  unsigned long long blocks = 0, plane_bytes = 0, stream_bytes = 0;
  unsigned int max_scan_bytes = 0;
  int max_pix = 0;
  for (int i = 0; i < n_images; ++i) {
    const JpegImage& m = im[i];
    JpegGeom& g = geom[(size_t)i];
    if (m.width == 0 || m.height == 0) { snprintf(g_jerr, sizeof(g_jerr), "image %d has zero size", i); return g_jerr; }
    if (!((m.n_comp == 1 && m.sampling == 0) || (m.n_comp == 3 && (m.sampling == 1 || m.sampling == 2)))) {
      snprintf(g_jerr, sizeof(g_jerr), "image %d: unsupported component layout (n_comp %d, sampling %d)", i, m.n_comp, m.sampling);
      return g_jerr;
    }
    for (int c = 0; c < m.n_comp; ++c)
      if (m.qt[c] >= n_q || m.dc[c] >= n_h || m.ac[c] >= n_h) {
        snprintf(g_jerr, sizeof(g_jerr), "image %d: table index out of range", i);
        return g_jerr;
      }
    g.ustream_offset = stream_bytes;
    stream_bytes += ((unsigned long long)m.scan_bytes + 32 + 3) & ~3ull;
    if (m.scan_bytes > max_scan_bytes) max_scan_bytes = m.scan_bytes;
    const int mcu = m.sampling == 2 ? 16 : 8;
    g.mcux = (m.width + mcu - 1) / mcu;
    g.mcuy = (m.height + mcu - 1) / mcu;
    const int f = m.sampling == 2 ? 2 : 1;
    g.bw[0] = g.mcux * f; g.bh[0] = g.mcuy * f;
    g.bw[1] = g.bw[2] = m.n_comp == 3 ? g.mcux : 0;
    g.bh[1] = g.bh[2] = m.n_comp == 3 ? g.mcuy : 0;
    g.coef_block0 = blocks;
    g.plane_offset = plane_bytes;
    const unsigned long long nb = (unsigned long long)g.bw[0] * g.bh[0] + 2ull * g.bw[1] * g.bh[1];
    blocks += nb;
    if (m.n_comp == 3) {
      plane_bytes += nb * 64;
      color_ids.push_back(i);
      if ((int)m.width * (int)m.height > max_pix) max_pix = (int)m.width * (int)m.height;
    }
  }
  if (color_ids.size() > 65535) return "more than 65535 colour images in one call";
  if (n_h > kJpegMaxTables) {
    snprintf(g_jerr, sizeof(g_jerr), "%d distinct Huffman tables in one call (at most %d; decode in smaller batches)", n_h, kJpegMaxTables);
    return g_jerr;
  }
  short* d_coef = nullptr;
  unsigned char* d_planes = nullptr;
  JpegImage* d_im = nullptr;
  JpegGeom* d_geom = nullptr;
  JpegHuff* d_h = nullptr;
  unsigned short* d_q = nullptr;
  int* d_cid = nullptr;
  cudaError_t e = cudaSuccess;
#define JCK(x) if ((e = (x)) != cudaSuccess) { snprintf(g_jerr, sizeof(g_jerr), "%s: %s", #x, cudaGetErrorString(e)); return g_jerr; }
  JCK(cudaMallocAsync(&d_coef, blocks * 64 * sizeof(short), st));
  JCK(cudaMemsetAsync(d_coef, 0, blocks * 64 * sizeof(short), st));
  if (plane_bytes) JCK(cudaMallocAsync(&d_planes, plane_bytes, st));
  // Descriptors go up in ONE copy from a pinned staging slot.  From pageable memory every cudaMemcpyAsync first
  // synchronises the stream: a loader that decodes ahead on a side stream would block the host -- and with it the
  // launches of the main stream -- until the previous decode has finished.
  const size_t sz_im = sizeof(JpegImage) * n_images, sz_geom = sizeof(JpegGeom) * n_images, sz_h = sizeof(JpegHuff) * n_h;
  const size_t sz_q = 128 * (size_t)n_q, sz_cid = sizeof(int) * color_ids.size();
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  const size_t o_im = 0, o_geom = al(sz_im), o_h = o_geom + al(sz_geom), o_q = o_h + al(sz_h), o_cid = o_q + al(sz_q);
  const size_t total = o_cid + al(sz_cid);
  PinnedSlot* slot = pinned_slot(total);
  if (!slot) return "pinned staging allocation failed";
  memcpy(slot->host + o_im, im, sz_im);
  memcpy(slot->host + o_geom, geom.data(), sz_geom);
  memcpy(slot->host + o_h, htables_host, sz_h);
  memcpy(slot->host + o_q, qtables_host, sz_q);
  if (sz_cid) memcpy(slot->host + o_cid, color_ids.data(), sz_cid);
  unsigned char* d_desc = nullptr;
  JCK(cudaMallocAsync(&d_desc, total, st));
  JCK(cudaMemcpyAsync(d_desc, slot->host, total, cudaMemcpyHostToDevice, st));
  JCK(cudaEventRecord(slot->done, st));
  d_im = reinterpret_cast<JpegImage*>(d_desc + o_im);
  d_geom = reinterpret_cast<JpegGeom*>(d_desc + o_geom);
  d_h = reinterpret_cast<JpegHuff*>(d_desc + o_h);
  d_q = reinterpret_cast<unsigned short*>(d_desc + o_q);
  d_cid = sz_cid ? reinterpret_cast<int*>(d_desc + o_cid) : nullptr;
  count_launch();
  // Images per warp: decoders sharing a warp diverge (every symbol costs the union of the lanes' paths: 500 frames take
  // 55 ms at 32 per warp, 23 ms at one per warp), but one image per warp oversubscribes the machine beyond ~16 warps
  // per SM (10 000 flow images: 115 ms vs 54 ms at 4 per warp).  Pick the smallest power of two that keeps the warp
  // count under 16 per SM.
  int per_warp = 1;
  while (per_warp < 32 && (n_images + per_warp - 1) / per_warp > 148 * 16) per_warp *= 2;
  if (const char* env = getenv("VA_JPEG_IMAGES_PER_WARP")) {
    const int v = atoi(env);
    if (v >= 1 && v <= 32 && (32 % v) == 0) per_warp = v;
  }
  const char* smem_env = getenv("VA_JPEG_SMEM_TABLES");
  const bool use_smem = !(smem_env != nullptr && smem_env[0] == '0');       // default: tables in shared memory
  HuffSmem* d_tab = nullptr;
  JCK(cudaMallocAsync(&d_tab, sizeof(HuffSmem), st));
  jpeg_tables_kernel<<<1, 256, 0, st>>>(d_h, n_h, d_tab);
  // Small batches of files without restart markers: one thread BLOCK per image (self-synchronising chunks); large batches
  // keep every SM busy with one thread per image, which does a quarter of the work per image.
  bool parallel = n_images <= 2048 && max_scan_bytes < (1u << 23);     // S <= 65536 bits: block counts fit 16 bits
  for (int i = 0; i < n_images && parallel; ++i) parallel = im[i].restart_interval == 0;
  if (const char* env = getenv("VA_JPEG_PARALLEL")) parallel = parallel && env[0] != '0';
  unsigned char* d_ustream = nullptr;
  if (parallel) {
    JCK(cudaMallocAsync(&d_ustream, stream_bytes + 64, st));
    jpeg_huffman_parallel_kernel<<<n_images, kParThreads, 0, st>>>(bitstreams, d_im, d_geom, d_tab, d_ustream, d_coef);
  } else if (use_smem) {
    jpeg_huffman_kernel<true><<<(n_images + per_warp - 1) / per_warp, 32, 0, st>>>(bitstreams, d_im, d_geom, d_h, d_tab, n_h,
                                                                                 n_images, 32 / per_warp, d_coef);
  } else {
    jpeg_huffman_kernel<false><<<(n_images + per_warp - 1) / per_warp, 32, 0, st>>>(bitstreams, d_im, d_geom, d_h, d_tab, n_h,
                                                                                  n_images, 32 / per_warp, d_coef);
  }
  if (d_ustream) cudaFreeAsync(d_ustream, st);
  cudaFreeAsync(d_tab, st);
  JCK(cudaGetLastError());
  count_launch();
  jpeg_idct_kernel<<<(unsigned)((blocks + 127) / 128), 128, 0, st>>>(d_coef, d_im, d_geom, d_q, n_images, blocks, d_planes, out);
  JCK(cudaGetLastError());
  if (!color_ids.empty()) {
    count_launch();
    jpeg_color_kernel<<<dim3((max_pix + 255) / 256, (unsigned)color_ids.size()), 256, 0, st>>>(d_cid, d_im, d_geom, d_planes, out);
    JCK(cudaGetLastError());
  }
  cudaFreeAsync(d_coef, st);
  if (d_planes) cudaFreeAsync(d_planes, st);
  cudaFreeAsync(d_desc, st);
#undef JCK
  return nullptr;
}

}  // namespace va
