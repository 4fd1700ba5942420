// Hotword ContextGraph (Aho-Corasick) for the device-resident beam search.
// Host build follows /root/reference core/hotword_context.py:34-137 literally (token_score = the phrase's full
// score on every edge, max() on shared prefixes without propagating to earlier-inserted descendants, output
// links accumulate output_score); the automaton is then flattened (BFS order, edges sorted by token id) into
// the arrays of ContextGraphDev. forward_one_step/finalize (:139-184) are one inline function used by both the
// CUDA search kernel and the host-side test entry points.
#pragma once
#include <stdint.h>

#include <vector>

#ifdef __CUDACC__
#define CG_HD __host__ __device__ __forceinline__
#else
#define CG_HD inline
#endif

namespace b200asr {

struct ContextGraphView {   // pointers valid on the side (host or device) that uses them
  int n_nodes;
  const int *edge_start, *edge_token, *edge_child, *fail, *token, *is_end, *output;
  const double *token_score, *node_score, *output_score;
};

CG_HD int cg_find_edge(const ContextGraphView &g, int node, int tok) {
  int lo = g.edge_start[node], hi = g.edge_start[node + 1] - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int t = g.edge_token[mid];
    if (t == tok) return g.edge_child[mid];
    if (t < tok) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// ContextGraph.forward_one_step, non-strict mode. Returns the score delta; *next = new state.
CG_HD double cg_forward_one_step(const ContextGraphView &g, int state, int tok, int *next) {
  int node;
  double score;
  const int direct = cg_find_edge(g, state, tok);
  if (direct >= 0) {
    node = direct;
    score = g.token_score[node];
  } else {
    node = g.fail[state];
    int child = cg_find_edge(g, node, tok);
    while (child < 0) {
      node = g.fail[node];
      if (g.token[node] == -1) { child = cg_find_edge(g, node, tok); break; }
      child = cg_find_edge(g, node, tok);
    }
    if (child >= 0) node = child;
    score = g.node_score[node] - g.node_score[state];
  }
  if (g.output_score[node] != 0) {
    double matched;
    if (g.is_end[node]) matched = g.node_score[node];
    else if (g.output[node] >= 0) matched = g.node_score[g.output[node]];
    else matched = g.node_score[node];
    *next = 0;
    return score + matched - g.node_score[node];
  }
  *next = node;
  return score;
}

CG_HD double cg_finalize(const ContextGraphView &g, int state) { return -g.node_score[state]; }

struct ContextGraphHost {
  std::vector<int> edge_start, edge_token, edge_child, fail, token, is_end, output;
  std::vector<double> token_score, node_score, output_score;
  int n_phrases = 0;
  int n_nodes() const { return (int)token.size(); }
  ContextGraphView view() const {
    return ContextGraphView{n_nodes(), edge_start.data(), edge_token.data(), edge_child.data(), fail.data(), token.data(),
                            is_end.data(), output.data(), token_score.data(), node_score.data(), output_score.data()};
  }
  // phrase p = tokens[offsets[p] .. offsets[p+1])
  void build(const int32_t *tokens, const int32_t *offsets, const float *scores, int n_phrases);
};

}  // namespace b200asr
