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

uint8_t dyn_smem[256 * 1024];

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
static pthread_barrier_t g_block;

void barrier () { g_sync.wait (); }

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

void launch (dim3 grid, dim3 block, size_t, const std::function<void ()> & body)
{
  const unsigned nt = block.x * block.y * block.z;
  std::vector<pthread_t> th (nt);
  std::vector<Job> jobs (nt);
  pthread_attr_t attr;
  pthread_attr_init (&attr);
  pthread_attr_setstacksize (&attr, 256 * 1024);
  pthread_barrier_init (&g_block, nullptr, nt);
  g_sync.reset (nt);
  for (unsigned t = 0; t < nt; t++) {
    jobs[t] = Job {&body, grid, block, t, nt};
    pthread_create (&th[t], &attr, thread_main, &jobs[t]);
  }
  for (unsigned t = 0; t < nt; t++) pthread_join (th[t], nullptr);
  pthread_barrier_destroy (&g_block);
  pthread_attr_destroy (&attr);
}

}  // namespace b200emu
