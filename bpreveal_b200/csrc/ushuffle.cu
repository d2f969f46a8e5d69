// k-let preserving shuffles of a one-hot sequence on the host (internal/libushuffle.c:59-344 as
// driven by ushuffle.shuffleOHE, ushuffle.py:42-69, from interpretUtils.py:1223-1225): the
// references of DeepSHAP when kmer-size > 1.  The reference runs this on the host as well.
//
// Algorithm (Jiang et al., uShuffle): the sequence is an Euler trail through the multigraph whose
// vertices are its distinct (k-1)-lets; a shuffle is a random arborescence towards the last let
// (Wilson's loop-erased walks) plus a random order of the remaining out-edges, walked from the
// first let.  To reproduce the reference's shuffles bit for bit, random numbers are drawn from a
// restatement of glibc's random() (TYPE_3 additive feedback, r[i] = r[i-3] + r[i-31], seeded by the
// 16807 Lehmer sequence, first 310 outputs discarded) in exactly the reference's order: vertices
// are numbered by first occurrence, edge lists keep their permutation from one shuffle to the
// next, Fisher-Yates runs from the top index down with `random() % (i + 1)`.
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace bp {

class GlibcRandom {
 public:
  explicit GlibcRandom(unsigned seed) {
    int32_t r[34];
    r[0] = static_cast<int32_t>(seed == 0 ? 1u : seed);
    for (int i = 1; i < 31; ++i) {
      // 16807 * r mod (2^31 - 1) by Schrage's decomposition, as srandom_r does
      const int32_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
      int32_t w = 16807 * lo - 2836 * hi;
      if (w < 0) w += 2147483647;
      r[i] = w;
    }
    for (int i = 0; i < 31; ++i) state_[i] = static_cast<uint32_t>(r[i]);
    front_ = 3;
    rear_ = 0;
    for (int i = 0; i < 310; ++i) next();
  }
  long next() {
    state_[front_] += state_[rear_];
    const uint32_t out = state_[front_] >> 1;
    front_ = (front_ + 1 == 31) ? 0 : front_ + 1;
    rear_ = (rear_ + 1 == 31) ? 0 : rear_ + 1;
    return static_cast<long>(out);
  }

 private:
  uint32_t state_[31];
  int front_, rear_;
};

struct KletGraph {
  int n_vertices = 0, root = 0;
  std::vector<int> first_pos;   // vertex -> position of its first occurrence in the sequence
  std::vector<int> edge_begin;  // vertex -> offset of its out-edge list (size n_vertices + 1)
  std::vector<int> edges;       // out-edges (target vertices) in sequence order, then as permuted
};

static KletGraph build_graph(const std::string& s, int k) {
  KletGraph g;
  const int l = static_cast<int>(s.size());
  const int n_lets = l - k + 2;  // number of (k-1)-lets
  std::unordered_map<std::string, int> ids;
  std::vector<int> let_vertex(n_lets);
  for (int i = 0; i < n_lets; ++i) {
    auto it = ids.emplace(s.substr(i, k - 1), g.n_vertices);
    if (it.second) {
      g.first_pos.push_back(i);
      ++g.n_vertices;
    }
    let_vertex[i] = it.first->second;
  }
  g.root = let_vertex[n_lets - 1];
  std::vector<int> deg(g.n_vertices, 0);
  for (int i = 0; i + 1 < n_lets; ++i) ++deg[let_vertex[i]];
  g.edge_begin.assign(g.n_vertices + 1, 0);
  for (int v = 0; v < g.n_vertices; ++v) g.edge_begin[v + 1] = g.edge_begin[v] + deg[v];
  g.edges.assign(n_lets - 1, 0);
  std::vector<int> fill(g.n_vertices, 0);
  for (int i = 0; i + 1 < n_lets; ++i) {
    const int u = let_vertex[i];
    g.edges[g.edge_begin[u] + fill[u]++] = let_vertex[i + 1];
  }
  return g;
}

template <typename T>
static void fisher_yates(T* t, int l, GlibcRandom& rng) {
  for (int i = l - 1; i > 0; --i) {
    const int j = static_cast<int>(rng.next() % (i + 1));
    const T tmp = t[i];
    t[i] = t[j];
    t[j] = tmp;
  }
}

static void one_shuffle(const std::string& s, int k, KletGraph& g, std::vector<int>& next,
                        GlibcRandom& rng, char* t) {
  const int l = static_cast<int>(s.size());
  if (k >= l) {
    s.copy(t, l);
    return;
  }
  if (k <= 1) {
    s.copy(t, l);
    fisher_yates(t, l, rng);
    return;
  }
  const int nv = g.n_vertices;
  std::vector<char> intree(nv, 0);
  intree[g.root] = 1;
  for (int i = 0; i < nv; ++i) {  // Wilson: loop-erased random walks into the tree
    int u = i;
    while (!intree[u]) {
      next[u] = static_cast<int>(rng.next() % (g.edge_begin[u + 1] - g.edge_begin[u]));
      u = g.edges[g.edge_begin[u] + next[u]];
    }
    u = i;
    while (!intree[u]) {
      intree[u] = 1;
      u = g.edges[g.edge_begin[u] + next[u]];
    }
  }
  for (int i = 0; i < nv; ++i) {  // tree edge last, the others in random order
    int* e = g.edges.data() + g.edge_begin[i];
    const int n = g.edge_begin[i + 1] - g.edge_begin[i];
    if (i != g.root) {
      const int j = e[n - 1];
      e[n - 1] = e[next[i]];
      e[next[i]] = j;
      fisher_yates(e, n - 1, rng);
    } else {
      fisher_yates(e, n, rng);
    }
  }
  s.copy(t, k - 1);  // the first let stays
  std::vector<int> used(nv, 0);
  int u = 0, i = k - 1;
  while (used[u] < g.edge_begin[u + 1] - g.edge_begin[u]) {
    const int v = g.edges[g.edge_begin[u] + used[u]];
    t[i++] = s[g.first_pos[v] + k - 2];
    ++used[u];
    u = v;
  }
}

}  // namespace bp

// host_in [length][alphabet] (any non-zero byte is "hot"), host_out [numShuffles][length][alphabet]
// of 0/1 bytes.  The generator is seeded with `seed` at the start of the call, which is what
// _ShapBatcher does for every query (seed 355687).
extern "C" int bpb_ushuffle_ohe(const unsigned char* host_in, unsigned char* host_out, int alphabet,
                                int length, int kmerSize, int numShuffles, unsigned seed) {
  using namespace bp;
  BP_REQUIRE(host_in && host_out, "bpb_ushuffle_ohe: NULL argument");
  BP_REQUIRE(alphabet >= 1 && alphabet <= 8 && length >= 1 && numShuffles >= 0,
             "bpb_ushuffle_ohe: alphabet must be 1..8, length >= 1");
  std::string s(static_cast<size_t>(length), '\0');
  for (int p = 0; p < length; ++p) {
    unsigned c = 0;
    for (int a = 0; a < alphabet; ++a) c |= (host_in[static_cast<size_t>(p) * alphabet + a] ? 1u : 0u) << a;
    s[p] = static_cast<char>(c);
  }
  GlibcRandom rng(seed);
  KletGraph g;
  std::vector<int> next;
  if (kmerSize > 1 && kmerSize < length) {
    g = build_graph(s, kmerSize);
    next.assign(g.n_vertices, 0);
  }
  std::vector<char> t(static_cast<size_t>(length));
  for (int n = 0; n < numShuffles; ++n) {
    one_shuffle(s, kmerSize, g, next, rng, t.data());
    unsigned char* o = host_out + static_cast<size_t>(n) * length * alphabet;
    for (int p = 0; p < length; ++p)
      for (int a = 0; a < alphabet; ++a)
        o[static_cast<size_t>(p) * alphabet + a] = (static_cast<unsigned char>(t[p]) >> a) & 1u;
  }
  return 0;
}
