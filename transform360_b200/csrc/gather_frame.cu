// sm_100a kernels of the projection-remap hot path, part 2: the persistent frame gather.
//
//   gatherFrameKernel<K, COPIES, GROUPS>   replaces cv::remap as the reference calls it (VideoFrameTransform.cpp:748-754)
//   for all planes of a frame in ONE launch: per output pixel a K x K window of the 8-bit source is weighted with
//   OpenCV's 15-bit fixed-point table and rounded with (sum + 16384) >> 15.  Bit-exact by construction: same table
//   (host-built, sampling.cpp), same integer arithmetic.
//
// This is a gather, not a contraction: no tensor cores.  What the design is built around (formats: kernels.cuh):
//   * One CTA per SM, GROUPS independent job pipelines of 8 warps each, ONE weight-table image in shared memory for
//     all of them (brought in by cp.async.bulk), so that the cubic table can be kept twice (bank-group balancing by
//     the host: 4.2-4.5 wavefronts per 128-bit weight load instead of 6.3-7.5) and Lanczos4's 128 KB table serves 24
//     warps instead of 16.
//   * The source window of a job arrives by ONE cp.async.bulk.tensor.2d (TMA) box load from the pitch-linear plane,
//     double-buffered against the arithmetic through mbarriers; taps are read as aligned 32-bit shared-memory words and
//     aligned with a funnel shift; every window row is folded with IDP.2A (two s16 x u8 multiply-adds per instruction).
//   * Share jobs (64 x 32 pixels, 2/3 of a cube map): a thread slides one K-row register window down its output
//     column and fetches only the 0-2 new source rows per pixel; 2.5 bytes of plan per pixel.
//   * Other staged jobs (32 x 32): a window per pixel, 4 bytes of plan per pixel; general jobs (pole caps, anything
//     BORDER_WRAP touches vertically) read their taps through L1 inside the same launch.
//   * Jobs are handed out by an atomic counter that re-arms itself; programmatic dependent launch lets the next
//     frame's prologue run under this frame's tail.
#include "gather_common.cuh"

#include <cuda.h>  // CUtensorMap (type only; no libcuda symbol is referenced)

#include <algorithm>
#include <cstring>
#include <utility>

namespace t360 {

namespace {

template <class F, int... I>
__device__ __forceinline__ void staticForImpl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void staticFor(F&& f) {  // f(integral_constant<int, 0>) ... f(integral_constant<int, N-1>)
  staticForImpl(f, std::make_integer_sequence<int, N>{});
}

// shared-memory loads with immediate offsets (one address register per pixel / per column)
template <int IMM>
__device__ __forceinline__ uint32_t ldsWordImm(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ uint4 ldsVecImm(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ uint2 ldsVec2Imm(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(IMM));
  return v;
}

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(arrivals));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmaLoadBox(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smemAddr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void bulkCopyToShared(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void groupBarrier(int group) {  // the 8 warps of one job pipeline
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
}
// compact records are streamed once per frame: read-only path, do not allocate in L1
__device__ __forceinline__ uint4 loadRecords128(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t loadRecords32(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// ---- window arithmetic -------------------------------------------------------------------------------------------
template <int K>
struct RowBytes {
  uint32_t b[K == 8 ? 2 : 1];  // the K source bytes of one window row, already aligned (K = 2: the low two bytes)
};

// row at byte offset IMM from the (word-aligned) address; sh = 8 * (byte offset of the window inside its first word)
template <int K, int IMM>
__device__ __forceinline__ RowBytes<K> loadRow(uint32_t addr, int sh) {
  RowBytes<K> o;
  const uint32_t q0 = ldsWordImm<IMM>(addr), q1 = ldsWordImm<IMM + 4>(addr);
  o.b[0] = __funnelshift_r(q0, q1, sh);
  if constexpr (K == 8) o.b[1] = __funnelshift_r(q1, ldsWordImm<IMM + 8>(addr), sh);
  return o;
}

// sum + 16384 over the window; slotAddr = shared address of the slot's first weight vector (VS bytes between vectors)
template <int K, int VS>
__device__ __forceinline__ int foldRows(const RowBytes<K> (&W)[K], uint32_t slotAddr) {
  int acc = 1 << 14;  // the rounding constant of FixedPtCast<int, uchar, 15>
  if constexpr (K == 2) {
    const uint2 w = ldsVec2Imm<0>(slotAddr);
    acc = dp2aLo(w.x, W[0].b[0], acc);
    acc = dp2aLo(w.y, W[1].b[0], acc);
  } else if constexpr (K == 4) {
    const uint4 wa = ldsVecImm<0>(slotAddr), wb = ldsVecImm<VS>(slotAddr);
    acc = dp2aLo(wa.x, W[0].b[0], acc); acc = dp2aHi(wa.y, W[0].b[0], acc);
    acc = dp2aLo(wa.z, W[1].b[0], acc); acc = dp2aHi(wa.w, W[1].b[0], acc);
    acc = dp2aLo(wb.x, W[2].b[0], acc); acc = dp2aHi(wb.y, W[2].b[0], acc);
    acc = dp2aLo(wb.z, W[3].b[0], acc); acc = dp2aHi(wb.w, W[3].b[0], acc);
  } else {
    staticFor<8>([&](auto R) {
      constexpr int r = decltype(R)::value;
      const uint4 wt = ldsVecImm<r * VS>(slotAddr);
      acc = dp2aLo(wt.x, W[r].b[0], acc); acc = dp2aHi(wt.y, W[r].b[0], acc);
      acc = dp2aLo(wt.z, W[r].b[1], acc); acc = dp2aHi(wt.w, W[r].b[1], acc);
    });
  }
  return acc;
}
__device__ __forceinline__ int biasedToByte(int acc) { return min(max(acc >> 15, 0), 255); }  // acc already holds + 16384

// K = 2 slots are 8 bytes: the record's slot field (slot << 4) is halved
template <int K>
__device__ __forceinline__ uint32_t slotOffset(uint32_t field) { return K == 2 ? field >> 1 : field; }

// ---- share job: one register window per output column, slid down 8 rows ---------------------------------------------
template <int K, int PITCH, int VS>
__device__ __forceinline__ void computeShareJob(const PlaneView& pv, uint32_t stageAddr, int outX, int outY, const uint4& rec,
                                                uint32_t header, uint32_t wAddr, int warp) {
  static_assert(PITCH % 4 == 0 && (K == 4 || K == 8), "");
  const int wx = warp & 1, wy = warp >> 1;
  const int dstPitch = pv.dstPitch;
  uint8_t* const dst = pv.dst + (size_t)(outY + wy * kShareRows) * dstPitch + (outX + wx * 32 + (int)(header >> kRecordColumnShift));
  uint32_t rowAddr = stageAddr + (header & 0x3ffcu);
  const int sh = (int)(header << 3);  // the funnel shift takes the low five bits: 8 * (offset & 3)
  RowBytes<K> W[K];
  staticFor<K>([&](auto R) { W[decltype(R)::value] = loadRow<K, decltype(R)::value * PITCH>(rowAddr, sh); });
  const uint32_t words[4] = {rec.x, rec.y, rec.z, rec.w};
  staticFor<kShareRows>([&](auto J) {
    constexpr int j = decltype(J)::value;
    const uint32_t r = (j & 1) ? words[j >> 1] >> 16 : words[j >> 1];
    if constexpr (j > 0) {
      const uint32_t d = r & 3u;  // source rows between this pixel's window and the previous one's: 0, 1 or 2
      if (d != 0) {
        rowAddr += d * PITCH;
        const RowBytes<K> last = loadRow<K, (K - 1) * PITCH>(rowAddr, sh);
        RowBytes<K> prev = W[K - 1];
        if (d == 2) prev = loadRow<K, (K - 2) * PITCH>(rowAddr, sh);
#pragma unroll
        for (int q = 0; q + 2 < K; ++q)
#pragma unroll
          for (int i = 0; i < (K == 8 ? 2 : 1); ++i) W[q].b[i] = d == 1 ? W[q + 1].b[i] : W[q + 2].b[i];
        W[K - 2] = prev;
        W[K - 1] = last;
      }
    }
    const int acc = foldRows<K, VS>(W, wAddr + (r & kSlotFieldMask));
    dst[(size_t)j * dstPitch] = (uint8_t)biasedToByte(acc);
  });
}

// ---- 32 x 32 staged job: a window per pixel, four pixels per thread ----------------------------------------------------
template <int K, int PITCH, int VS>
__device__ __forceinline__ void computeTileJob(const PlaneView& pv, uint32_t stageAddr, int outX, int outY, const uint4& rec,
                                               uint32_t wAddr, int warp) {
  static_assert(PITCH % 4 == 0, "");
  const int y0 = outY + warp * kRowsPerThread;
  const int dstPitch = pv.dstPitch, dstW = pv.dstW, dstH = pv.dstH;
  uint8_t* const dstRow = pv.dst + (size_t)y0 * dstPitch + outX;
  const uint32_t words[4] = {rec.x, rec.y, rec.z, rec.w};
  staticFor<kRowsPerThread>([&](auto J) {
    constexpr int j = decltype(J)::value;
    const uint32_t w = words[j];
    const int col = (int)(w >> 16) & 31;
    if (outX + col < dstW && y0 + j < dstH) {
      const uint32_t rowAddr = stageAddr + (w & 0x7ffcu);
      const int sh = (int)(w << 3);
      RowBytes<K> W[K];
      staticFor<K>([&](auto R) { W[decltype(R)::value] = loadRow<K, decltype(R)::value * PITCH>(rowAddr, sh); });
      const int acc = foldRows<K, VS>(W, wAddr + slotOffset<K>((w >> 17) & kSlotFieldMask));
      dstRow[(size_t)j * dstPitch + col] = (uint8_t)biasedToByte(acc);
    }
  });
}

struct FrameTensorMaps {
  CUtensorMap map[kMaxFramePlanes][kNumBoxClasses];
};

template <int K, int COPIES, int GROUPS>
__host__ __device__ constexpr int frameSmemBytes() {
  return weightImageBytes(K, COPIES) + GROUPS * 2 * stageBytesOf(K) + (GROUPS + 1) * 64;
}

template <int K, int COPIES, int GROUPS>
__global__ void __launch_bounds__(GROUPS * kGroupThreads, 1)
gatherFrameKernel(const __grid_constant__ FrameGatherParams p, StagedParams jobs, const __grid_constant__ FrameTensorMaps maps) {
  constexpr int VS = weightVectorStride(K, COPIES), kWeights = weightImageBytes(K, COPIES), kStage = stageBytesOf(K);
  constexpr uint32_t kBox0 = stageBoxW(K, 0) * stageBoxH(K, 0), kBox1 = stageBoxW(K, 1) * stageBoxH(K, 1),
                     kBoxShare = stageBoxW(K, 2) * stageBoxH(K, 2);
  static_assert(kBox1 + 64 <= 2 * kStage && kBox0 + 64 <= kStage && kBoxShare + 64 <= kStage, "boxes must fit their stage buffers");
  static_assert(kBox0 % 16 == 0 && kStage % 128 == 0 && kWeights % 128 == 0, "alignment of the seam merge / TMA destinations");
  extern __shared__ __align__(128) unsigned char smem[];
  const int group = threadIdx.x / kGroupThreads, t = threadIdx.x % kGroupThreads;
  const int lane = t & 31, warp = t >> 5;
  unsigned char* wsmem = smem;
  unsigned char* stage0 = smem + kWeights + group * (2 * kStage);
  uint64_t* barBase = reinterpret_cast<uint64_t*>(smem + kWeights + GROUPS * 2 * kStage);
  uint64_t* bars = barBase + group * 8;  // [0], [1]: stage buffers, [2]: both buffers together, [3]: two claim slots
  uint64_t* weightBar = barBase + GROUPS * 8;

  // Programmatic dependent launch: the next launch on the stream (the next frame's gather) may place its CTAs as soon
  // as ours retire, and run its prologue -- which touches only constant data: weights, job list, sampling records --
  // under our tail.  Everything an earlier kernel may have written (the source planes, the scheduler counters) is
  // only touched after griddepcontrol.wait below.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int g = 0; g < GROUPS; ++g)
      for (int i = 0; i < 3; ++i) mbarInit(barBase + g * 8 + i, 1);
    mbarInit(weightBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the weight image (host-permuted, both copies) in four bulk copies
    mbarExpectTx(weightBar, kWeights);
    constexpr int kChunk = kWeights / 4;
    for (int i = 0; i < 4; ++i)
      bulkCopyToShared(wsmem + i * kChunk, reinterpret_cast<const unsigned char*>(p.weightImage) + i * kChunk, kChunk, weightBar);
  }
  __syncthreads();

  const int worker = blockIdx.x * GROUPS + group, numWorkers = gridDim.x * GROUPS;
  // Software pipeline, two deep, so that no load is waited for in the iteration that issues it (a warp executes in
  // order: a header load followed by the record loads that need its fields would stall the whole job on the header):
  //   iteration n:  issue header(n+2) | issue records(n+1) from header(n+1), already in registers | compute job n
  auto loadHeader = [&](int i) { return i < jobs.numTiles ? jobs.tiles[i] : GatherJob{0, 0, 0, 0}; };
  // The header fetched two jobs ahead must not be waited for where it is issued.  The compiler keeps warp-uniform
  // values in uniform registers and converts a loaded header the moment it arrives; so the load is opaque (asm: four
  // ordinary registers), and the header becomes uniform -- through a warp reduction whose result the compiler knows
  // to be uniform -- only at the end of the job, when it has long arrived.  An index past the list reads its last
  // entry; validity is tracked by the index itself.
  auto issueHeaderLoad = [&](int i, int (&raw)[4]) {
    const GatherJob* src = jobs.tiles + min(i, jobs.numTiles - 1);
    asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(raw[0]), "=r"(raw[1]), "=r"(raw[2]), "=r"(raw[3]) : "l"(src));
  };
  auto uniformHeader = [&](const int (&raw)[4]) {
    return GatherJob{(int)__reduce_or_sync(0xffffffffu, (unsigned)raw[0]), (int)__reduce_or_sync(0xffffffffu, (unsigned)raw[1]),
                     (int)__reduce_or_sync(0xffffffffu, (unsigned)raw[2]), (int)__reduce_or_sync(0xffffffffu, (unsigned)raw[3])};
  };
  // The plane of a job, field by field through selects on kernel-parameter operands: an indexed load from the
  // parameter bank instead would put its latency in front of every job's record loads.
  static_assert(kMaxFramePlanes == 3, "planeOf selects among three planes");
  const PlaneView &pa = p.plane[0], &pb = p.plane[1], &pc = p.plane[2];
#define T360_PICK(pl, f) ((pl) == 0 ? pa.f : ((pl) == 1 ? pb.f : pc.f))
  auto planeOf = [&](const GatherJob& job) {
    const int pl = job.outY >> kJobPlaneShift;
    return PlaneView{T360_PICK(pl, src), T360_PICK(pl, dst), T360_PICK(pl, samples), nullptr, T360_PICK(pl, srcW), T360_PICK(pl, srcH),
                     T360_PICK(pl, srcPitch), T360_PICK(pl, dstW), T360_PICK(pl, dstH), T360_PICK(pl, dstPitch), T360_PICK(pl, tilesPerRow), 0};
  };
  auto kindOf = [](const GatherJob& job) { return (job.outY >> kJobKindShift) & kJobKindMask; };
  // compact records: one 128-bit load per thread (+ one 32-bit column header in a share job); general jobs fetch their
  // full records when they run (few jobs, latency-bound anyway)
  auto loadRecords = [&](int i, const GatherJob& job, uint4& rec, uint32_t& header) {
    rec = make_uint4(0, 0, 0, 0);
    header = 0;
    const int kind = kindOf(job);
    if (i >= jobs.numTiles || kind == kJobGeneral) return;
    const int pl = job.outY >> kJobPlaneShift;
    const uint4* base = T360_PICK(pl, records) + (unsigned)job.recordOffset;
    if (kind == kJobShare) {
      base += warp * (kShareJobRecordBytes / kGroupWarps / 16);
      rec = loadRecords128(base + lane);
      header = loadRecords32(reinterpret_cast<const uint32_t*>(base + 32) + lane);
    } else {
      rec = loadRecords128(base + warp * 32 + lane);
    }
  };
  // Dynamic scheduling: the first four jobs of a group are static (worker + k * numWorkers), every further one is
  // claimed from a global counter by the group's thread 0 and handed to the other threads through a double-buffered
  // shared slot across the end-of-job barrier.  The value the atomic returns is not touched in the iteration that
  // issues it -- a warp executes in order and would sit out the round trip while the rest of the group waits for it at
  // the barrier -- but one iteration later (in an asm statement, so that the compiler cannot hoist the use).
  int* claimSlot = reinterpret_cast<int*>(bars + 3);
  const int claimBase = 4 * numWorkers;
  int claimedRaw = worker - numWorkers;  // thread 0; claimBase + claimedRaw = the group's fourth static job
  int i0 = worker, i1 = i0 + numWorkers, i2 = i1 + numWorkers;
  GatherJob job = loadHeader(i0), jobNext = loadHeader(i1);
  uint4 rec;
  uint32_t header;
  loadRecords(i0, job, rec, header);
  // q0 / q1: single-buffer boxes (class 0, share) / double-buffer boxes (class 1, seam) this group has consumed;
  // issued0: single-buffer boxes it has requested.  A single-buffer job with sequence number q lives in stage q & 1 and
  // completes phase (q >> 1) & 1 of that stage's barrier.
  uint32_t q0 = 0, q1 = 0, issued0 = 0;
  const uint32_t wAddr = smemAddr(wsmem);
  mbarWait(weightBar, 0);
  asm volatile("griddepcontrol.wait;" ::: "memory");  // earlier kernels on the stream are complete and visible from here on
  auto pipelined = [](int kind) { return kind == kJobClass0 || kind == kJobShare; };
  auto requestBox = [&](const GatherJob& j) {  // thread 0 of the group only
    const uint32_t st = issued0 & 1;
    const bool share = kindOf(j) == kJobShare;
    mbarExpectTx(&bars[st], share ? kBoxShare : kBox0);
    tmaLoadBox(stage0 + st * kStage, &maps.map[j.outY >> kJobPlaneShift][share ? 2 : 0], j.boxXY & 0xffff, j.boxXY >> 16, &bars[st]);
  };
  for (uint32_t it = 0; i0 < jobs.numTiles; ++it) {
    const int next = i1;
    if (t == 0) {
      int claimed;
      asm volatile("add.s32 %0, %1, %2;" : "=r"(claimed) : "r"(claimedRaw), "r"(claimBase));
      claimSlot[it & 1] = claimed;
      claimedRaw = atomicAdd(jobs.claimCounter, 1);
    }
    int headerAfterNext[4];
    issueHeaderLoad(i2, headerAfterNext);
    uint4 recNext;
    uint32_t headerNext;
    loadRecords(next, jobNext, recNext, headerNext);
    const int kind = kindOf(job), outY = job.outY & kJobRowMask;
    const PlaneView pv = planeOf(job);
    const bool nextPipelined = next < jobs.numTiles && pipelined(kindOf(jobNext));
    if (pipelined(kind)) {
      if (issued0 == q0) {  // not prefetched (first job, or it follows a job that needed both stages)
        if (t == 0) requestBox(job);
        ++issued0;
      }
      if (nextPipelined) {  // the other stage was released by the barrier that ended the previous job
        if (t == 0) requestBox(jobNext);
        ++issued0;
      }
      const uint32_t st = q0 & 1;
      mbarWait(&bars[st], (q0 >> 1) & 1);
      const uint32_t stageAddr = smemAddr(stage0 + st * kStage);
      if (kind == kJobShare) {
        if constexpr (K >= 4) computeShareJob<K, stageBoxW(K, 2), VS>(pv, stageAddr, job.outX, outY, rec, header, wAddr, warp);
      } else {
        computeTileJob<K, stageBoxW(K, 0), VS>(pv, stageAddr, job.outX, outY, rec, wAddr, warp);
      }
      ++q0;
    } else if (kind == kJobClass1) {
      if (t == 0) {  // no single-buffer box is in flight here: the larger box may span both stage buffers
        mbarExpectTx(&bars[2], kBox1);
        tmaLoadBox(stage0, &maps.map[job.outY >> kJobPlaneShift][1], job.boxXY & 0xffff, job.boxXY >> 16, &bars[2]);
      }
      mbarWait(&bars[2], q1 & 1);
      computeTileJob<K, stageBoxW(K, 1), VS>(pv, smemAddr(stage0), job.outX, outY, rec, wAddr, warp);
      ++q1;
    } else if (kind == kJobSeam) {
      // two complementary class-0 boxes (zero-filled outside the plane), one per stage buffer, OR-ed into the first
      const int boxX = job.boxXY & 0xffff, boxY = job.boxXY >> 16;
      if (t == 0) {
        mbarExpectTx(&bars[2], 2 * kBox0);
        const CUtensorMap* m = &maps.map[job.outY >> kJobPlaneShift][0];
        tmaLoadBox(stage0, m, boxX, boxY, &bars[2]);
        tmaLoadBox(stage0 + kStage, m, boxX - pv.srcW, boxY, &bars[2]);
      }
      mbarWait(&bars[2], q1 & 1);
      {
        uint4* a = reinterpret_cast<uint4*>(stage0);
        const uint4* b = reinterpret_cast<const uint4*>(stage0 + kStage);
        for (int i = t; i < (int)(kBox0 / 16); i += kGroupThreads) {
          uint4 x = a[i];
          const uint4 y = b[i];
          x.x |= y.x; x.y |= y.y; x.z |= y.z; x.w |= y.w;
          a[i] = x;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // these writes precede later TMA writes to the stage
      }
      groupBarrier(group);
      computeTileJob<K, stageBoxW(K, 0), VS>(pv, smemAddr(stage0), job.outX, outY, rec, wAddr, warp);
      ++q1;
    } else {
      if (nextPipelined && issued0 == q0) {  // both stages are idle during a general job: start the next box now
        if (t == 0) requestBox(jobNext);
        ++issued0;
      }
      SrcView sv;
      sv.bytes = pv.src;
      sv.misalign = (int)(reinterpret_cast<uintptr_t>(pv.src) & 3);
      sv.words = reinterpret_cast<const uint32_t*>(pv.src - sv.misalign);
      sv.w = pv.srcW; sv.h = pv.srcH; sv.pitch = pv.srcPitch;
      const int y0 = outY + warp * kRowsPerThread;
      if (job.outX + lane < pv.dstW) {
        // full records, tile-major over tiles of 32 x gatherTileH(K) pixels
        const int2* segment = pv.samples + ((size_t)(y0 / gatherTileH(K)) * pv.tilesPerRow + job.outX / kGatherTileW) * gatherTileH(K) * kGatherTileW +
                              (y0 % gatherTileH(K)) * kGatherTileW + lane;
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j) {
          if (y0 + j >= pv.dstH) break;
          const int2 full = loadPlan(segment + j * kGatherTileW);
          const int v = gatherPixel<K, false, VS>(sv, wsmem, recordCol0(full.x), full.y);
          pv.dst[(size_t)(y0 + j) * pv.dstPitch + job.outX + recordColumn(full.x)] = (uint8_t)v;
        }
      }
    }
    groupBarrier(group);  // everyone is done with this job's stage before it is refilled (and sees the claimed index)
    i0 = i1; i1 = i2; i2 = claimSlot[it & 1];
    job = jobNext;
    jobNext = uniformHeader(headerAfterNext);
    // (asm: the copies stay here, ahead of the next job's loads, whose scoreboards they would otherwise share)
    asm volatile("mov.b32 %0, %1;" : "=r"(rec.x) : "r"(recNext.x));
    asm volatile("mov.b32 %0, %1;" : "=r"(rec.y) : "r"(recNext.y));
    asm volatile("mov.b32 %0, %1;" : "=r"(rec.z) : "r"(recNext.z));
    asm volatile("mov.b32 %0, %1;" : "=r"(rec.w) : "r"(recNext.w));
    asm volatile("mov.b32 %0, %1;" : "=r"(header) : "r"(headerNext));
  }
#undef T360_PICK
  // the group that finishes last re-arms the scheduler for the next launch (claimCounter[0] = claims, [1] = finished groups)
  if (t == 0 && atomicAdd(jobs.claimCounter + 1, 1) == numWorkers - 1) {
    jobs.claimCounter[0] = 0;
    jobs.claimCounter[1] = 0;
    __threadfence();
  }
}

template <int K, int COPIES, int GROUPS>
cudaError_t launchFrameK(const FrameGatherParams& p, const StagedParams& jobs, const FrameTensorMaps& maps, int numSMs,
                         cudaStream_t stream) {
  static DeviceLaunchCfg cfgs;
  constexpr int threads = GROUPS * kGroupThreads, smemBytes = frameSmemBytes<K, COPIES, GROUPS>();
  LaunchCfg cfg;
  cudaError_t err = prepare<gatherFrameKernel<K, COPIES, GROUPS>>(cfgs, threads, smemBytes, cfg);
  if (err != cudaSuccess) return err;
  const int grid = std::min(numSMs * cfg.perSM, (jobs.numTiles + GROUPS - 1) / GROUPS);  // persistent: one CTA per SM
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(threads);
  lc.dynamicSmemBytes = smemBytes;
  lc.stream = stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see griddepcontrol.* in the kernel
  attr.val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = &attr;
  lc.numAttrs = 1;
  err = cudaLaunchKernelEx(&lc, gatherFrameKernel<K, COPIES, GROUPS>, p, jobs, maps);
  gLaunches.fetch_add(1, std::memory_order_relaxed);
  return err;
}

}  // namespace

cudaError_t launchGatherFrame(const FrameGatherParams& p, const StagedParams& jobs, const void* tensorMaps, int numSMs,
                              cudaStream_t stream) {
  if (jobs.numTiles <= 0) return cudaSuccess;
  if (p.numPlanes < 1 || p.numPlanes > kMaxFramePlanes) return cudaErrorInvalidValue;
  FrameTensorMaps maps;
  std::memcpy(&maps, tensorMaps, sizeof(CUtensorMap) * kNumBoxClasses * p.numPlanes);
  for (int i = p.numPlanes; i < kMaxFramePlanes; ++i)  // unused entries: valid descriptors that no job refers to
    for (int c = 0; c < kNumBoxClasses; ++c) maps.map[i][c] = maps.map[0][c];
  switch (p.kernelSize) {
    case 2: return launchFrameK<2, weightCopies(2), kGatherGroups>(p, jobs, maps, numSMs, stream);
    case 4: return launchFrameK<4, weightCopies(4), kGatherGroups>(p, jobs, maps, numSMs, stream);
    case 8: return launchFrameK<8, weightCopies(8), kGatherGroups>(p, jobs, maps, numSMs, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace t360
