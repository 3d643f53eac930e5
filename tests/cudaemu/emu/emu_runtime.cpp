// tests/cudaemu/emu/emu_runtime.cpp — TEST INFRASTRUCTURE: block-by-block execution of an emulated launch.
//
// One set of threads per launch; every thread walks the blocks of the grid in order.  __syncthreads() is a barrier over
// the threads of the block that have not returned from the kernel yet (what the hardware does), a second barrier
// separates consecutive blocks (shared memory is reused).
#include "cuda_runtime.h"

#include <condition_variable>
#include <mutex>

thread_local uint3_emu threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace b200emu {

uint8_t *dyn_smem = nullptr;

struct LiveBarrier {
  std::mutex m;
  std::condition_variable cv;
  unsigned expected = 0, arrived = 0;
  unsigned long generation = 0;
  void reset (unsigned n) { expected = n; arrived = 0; }
  void wait ()
  {
    std::unique_lock<std::mutex> l (m);
    const unsigned long g = generation;
    if (++arrived >= expected) { arrived = 0; generation++; cv.notify_all (); return; }
    cv.wait (l, [&] { return generation != g; });
  }
  void leave ()                  // a thread returned from the kernel: it no longer takes part in __syncthreads
  {
    std::unique_lock<std::mutex> l (m);
    if (expected) expected--;
    if (expected && arrived >= expected) { arrived = 0; generation++; cv.notify_all (); }
  }
};

static LiveBarrier g_sync;
static LiveBarrier g_warp[64];                 // one per warp of the block
static unsigned g_wa[64][32][4], g_wb[64][32][2];
static pthread_barrier_t g_block;

void barrier () { g_sync.wait (); }

static unsigned g_wx[64][32];

unsigned lane_id ()
{
  return (threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31;
}

void warp_gather (unsigned v, unsigned out[32])
{
  const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const unsigned w = tid >> 5, lane = tid & 31;
  g_wx[w][lane] = v;
  g_warp[w].wait ();
  for (int i = 0; i < 32; i++) out[i] = g_wx[w][i];
  g_warp[w].wait ();
}

static void warp_mma (int d[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1, const int c[4],
    bool a_signed);

void warp_mma_u8s8 (int d[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1, const int c[4])
{
  warp_mma (d, a0, a1, a2, a3, b0, b1, c, false);
}

void warp_mma_s8u8 (int d[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1, const int c[4])
{
  warp_mma (d, a0, a1, a2, a3, b0, b1, c, true);
}

static void warp_mma (int d[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1, const int c[4],
    bool a_signed)
{
  const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const unsigned w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  g_wa[w][lane][0] = a0; g_wa[w][lane][1] = a1; g_wa[w][lane][2] = a2; g_wa[w][lane][3] = a3;
  g_wb[w][lane][0] = b0; g_wb[w][lane][1] = b1;
  g_warp[w].wait ();
  auto A = [&] (unsigned row, unsigned k) -> int {                 // u8
    const unsigned src = (row & 7) * 4 + ((k & 15) >> 2), reg = (row >> 3) + 2 * (k >> 4);
    const unsigned v = (g_wa[w][src][reg] >> (8 * (k & 3))) & 0xff;
    return a_signed ? (int) (int8_t) v : (int) v;
  };
  auto B = [&] (unsigned k, unsigned n) -> int {                   // s8
    const unsigned src = n * 4 + ((k & 15) >> 2), reg = k >> 4;
    const unsigned v = (g_wb[w][src][reg] >> (8 * (k & 3))) & 0xff;
    return a_signed ? (int) v : (int) (int8_t) v;                  // the operand that is not the taps holds u8 pixels
  };
  const unsigned rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
  for (int i = 0; i < 4; i++) {
    int acc = c[i];
    for (unsigned k = 0; k < 32; k++) acc += A (rows[i], k) * B (k, cols[i]);
    d[i] = acc;
  }
  g_warp[w].wait ();
}

struct Job {
  const std::function<void ()> *body;
  dim3 grid, block;
  unsigned tid, nt;
};

static void *thread_main (void *arg)
{
  Job *j = (Job *) arg;
  gridDim = j->grid; blockDim = j->block;
  threadIdx.x = j->tid % j->block.x;
  threadIdx.y = (j->tid / j->block.x) % j->block.y;
  threadIdx.z = j->tid / (j->block.x * j->block.y);
  for (unsigned bz = 0; bz < j->grid.z; bz++)
    for (unsigned by = 0; by < j->grid.y; by++)
      for (unsigned bx = 0; bx < j->grid.x; bx++) {
        blockIdx = uint3_emu {bx, by, bz};
        (*j->body) ();
        g_sync.leave ();
        if (pthread_barrier_wait (&g_block) == PTHREAD_BARRIER_SERIAL_THREAD)
          g_sync.reset (j->nt);                                    // every thread is between blocks here
        pthread_barrier_wait (&g_block);
      }
  return nullptr;
}

void launch (dim3 grid, dim3 block, size_t smem_bytes, const std::function<void ()> & body)
{
  const unsigned nt = block.x * block.y * block.z;
  // exactly the requested bytes (rounded to 16): an address sanitizer build then catches a kernel that runs past them
  uint8_t *smem = (uint8_t *) aligned_alloc (16, std::max<size_t> (16, (smem_bytes + 15) & ~(size_t) 15));
  dyn_smem = smem;
  std::vector<pthread_t> th (nt);
  std::vector<Job> jobs (nt);
  pthread_attr_t attr;
  pthread_attr_init (&attr);
  pthread_attr_setstacksize (&attr, 256 * 1024);
  pthread_barrier_init (&g_block, nullptr, nt);
  g_sync.reset (nt);
  for (unsigned w = 0; w < 64; w++) g_warp[w].reset (32);
  for (unsigned t = 0; t < nt; t++) {
    jobs[t] = Job {&body, grid, block, t, nt};
    pthread_create (&th[t], &attr, thread_main, &jobs[t]);
  }
  for (unsigned t = 0; t < nt; t++) pthread_join (th[t], nullptr);
  pthread_barrier_destroy (&g_block);
  pthread_attr_destroy (&attr);
  dyn_smem = nullptr;
  free (smem);
}

}  // namespace b200emu
