// planner.cu -- work items of the tile sweep: (block of group1 rows) x (chunk of group2), per
// rank.  Pure host logic (also exported as b200cv_plan_items for the CPU-side sharding tests).
#include "engine.h"

using namespace b200cv;

namespace b200cv {

int ensure_partials(b200cv_cvc *h, int need)
{
  if (need > h->partial_cap) {
    free_dev(h->d_partials);
    CUDA_OK(h, cudaMalloc(&h->d_partials, sizeof(double) * need));
    h->partial_cap = need;
  }
  return B200CV_OK;
}

// Work items of the tile sweep for one rank: (group1 block) x (group2 chunk).  Pure host
// logic (no CUDA): also exported as b200cv_plan_items for the CPU-side sharding tests.
void plan_items(bool self, int n1, int n2_in, int rank, int world, std::vector<SweepItem> &mine,
                int &chunk_out)
{
  int const n2 = self ? n1 : n2_in;
  int const nb1 = (n1 + SWEEP_IB - 1) / SWEEP_IB;
  // chunk length: the longest that still fills the machine.  CTAs run two per SM (register
  // bound), i.e. in waves of 296 per GPU; a chunk costs (tiles + ~1.5 tiles of fixed
  // work), the last wave is as long as a full one.  Pick the candidate with the least
  // estimated time; big problems end up at JCMAX, a 1k x 100k problem at exactly one wave.
  long long const slots = 148LL * 2;
  long long chunk = JCMAX;
  double best = 1e300;
  for (int cand = 32 * SWEEP_W; cand <= JCMAX; cand += 32) {
    long long const nch = (n2 + cand - 1) / cand;
    long long items = (long long)nb1 * nch;
    if (self) items = items / 2 + nb1;  // upper triangle
    long long const per_rank = (items + world - 1) / world;
    long long const waves = (per_rank + slots - 1) / slots;
    double const t = (double)waves * ((double)cand / 32.0 + 1.5);
    if (t < best * (1.0 - 1e-9) || (t <= best * (1.0 + 1e-9) && cand > chunk)) {
      best = t;
      chunk = cand;
    }
  }
  if (const char *forced = std::getenv("B200CV_CHUNK")) {
    // tuning experiments only: group2 atoms per work item (multiple of 32, <= JCMAX)
    long long const c = std::atoll(forced);
    if (c >= 32 && c <= JCMAX && c % 32 == 0) chunk = c;
  }
  chunk_out = (int)chunk;
  int const nch = (n2 + (int)chunk - 1) / (int)chunk;

  std::vector<SweepItem> all;
  all.reserve((size_t)nb1 * nch);
  for (int b = 0; b < nb1; b++) {
    int const i0 = b * SWEEP_IB;
    for (int c = 0; c < nch; c++) {
      int const j0 = c * (int)chunk;
      int const jn = std::min((int)chunk, n2 - j0);
      int diag = 0;
      if (self) {
        if (j0 + jn <= i0 + 1) continue;       // no j > i in this chunk
        diag = (j0 < i0 + SWEEP_IB) ? 1 : 0;   // overlaps the block's own rows
      }
      all.push_back(SweepItem{i0, j0, jn, diag});
    }
  }
  mine.clear();
  if (world == 1) {
    mine.swap(all);
  } else if (!self) {
    // equal contiguous ranges of the block-major item list: a rank owns whole group1 blocks
    // except at the two ends of its range, and the load is balanced to one item
    size_t const lo = all.size() * (size_t)rank / (size_t)world;
    size_t const hi = all.size() * (size_t)(rank + 1) / (size_t)world;
    mine.assign(all.begin() + lo, all.begin() + hi);
  } else {
    // cyclic over the flat list (balances the selfCoordNum triangle)
    for (size_t k = 0; k < all.size(); k++) {
      if ((int)(k % (size_t)world) == rank) mine.push_back(all[k]);
    }
  }
  // The kernel ends when its last CTA does: with several waves of equal items the SMs drain
  // one by one over the duration of a whole item (0.4 ms for 32 tiles -- 2 % of a step at 8
  // GPUs).  Cut the items of the last two waves into quarters so the drain is a quarter as long.
  if ((long long)mine.size() > 2 * slots && !std::getenv("B200CV_CHUNK")) {
    size_t const first_small = mine.size() - (size_t)(2 * slots);
    std::vector<SweepItem> tail;
    tail.reserve((size_t)(8 * slots));
    for (size_t k = first_small; k < mine.size(); k++) {
      SweepItem const it = mine[k];
      int const tiles = (it.jn + 31) / 32;
      int const per = std::max(4, (tiles + 3) / 4);  // tiles per piece
      for (int t0 = 0; t0 < tiles; t0 += per) {
        int const j0 = it.j0 + t0 * 32;
        int const jn = std::min(per * 32, it.j0 + it.jn - j0);
        int diag = 0;
        if (self) {
          if (j0 + jn <= it.i0 + 1) continue;     // no j > i in this piece
          diag = (j0 < it.i0 + SWEEP_IB) ? 1 : 0;
        }
        tail.push_back(SweepItem{it.i0, j0, jn, diag});
      }
    }
    mine.resize(first_small);
    mine.insert(mine.end(), tail.begin(), tail.end());
  }
}

int build_items(b200cv_cvc *h)
{
  h->epoch++;
  std::vector<SweepItem> mine;
  plan_items(h->self(), h->g[0].n, h->g[1].n, h->rank, h->world, mine, h->chunk);
  free_dev(h->d_items);
  h->nitems = (int)mine.size();
  if (h->nitems > 0) {
    CUDA_OK(h, cudaMalloc(&h->d_items, sizeof(SweepItem) * mine.size()));
    CUDA_OK(h, cudaMemcpy(h->d_items, mine.data(), sizeof(SweepItem) * mine.size(),
                          cudaMemcpyHostToDevice));
  }
  int rc = ensure_partials(h, std::max(h->nitems, 1) + PARTIALS_TAIL);
  if (rc) return rc;
  h->items_dirty = false;
  return B200CV_OK;
}

}  // namespace b200cv

extern "C" {

int b200cv_plan_items(int kind, int n1, int n2, int rank, int world, int *items, size_t capacity,
                      size_t *count, int *chunk)
{
  if (!count || world < 1 || rank < 0 || rank >= world || n1 < 1) return B200CV_INPUT_ERROR;
  std::vector<SweepItem> mine;
  int ch = 0;
  plan_items(kind == B200CV_SELFCOORDNUM, n1, n2, rank, world, mine, ch);
  *count = mine.size();
  if (chunk) *chunk = ch;
  if (items) {
    if (capacity < mine.size()) return B200CV_MEMORY_ERROR;
    std::memcpy(items, mine.data(), sizeof(SweepItem) * mine.size());
  }
  return B200CV_OK;
}

}  // extern "C"
