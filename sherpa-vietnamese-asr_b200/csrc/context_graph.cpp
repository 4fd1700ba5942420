// Host-side build of the hotword automaton; see context_graph.h for the reference lines followed.
#include "context_graph.h"

#include <algorithm>
#include <deque>
#include <utility>

namespace b200asr {

namespace {
struct BuildNode {
  int token = -1;
  double token_score = 0.0, node_score = 0.0, output_score = 0.0;
  bool is_end = false;
  std::vector<std::pair<int, int>> next;  // (token, node) in insertion order, like a Python dict
  int fail = 0, output = -1;
  int find(int tok) const {
    for (const auto &e : next)
      if (e.first == tok) return e.second;
    return -1;
  }
};
}  // namespace

void ContextGraphHost::build(const int32_t *tokens, const int32_t *offsets, const float *scores, int n) {
  std::vector<BuildNode> nodes(1);
  n_phrases = 0;
  for (int p = 0; p < n; ++p) {
    const int b = offsets[p], e = offsets[p + 1];
    if (e <= b) continue;                       // empty sequences are skipped (:56-57)
    const double sc = (double)scores[p];
    int cur = 0;
    for (int j = b; j < e; ++j) {
      const int tid = tokens[j];
      const bool last = (j == e - 1);
      int nxt = nodes[cur].find(tid);
      if (nxt < 0) {
        BuildNode c;
        c.token = tid;
        c.token_score = sc;
        c.node_score = nodes[cur].node_score + sc;
        c.output_score = last ? c.node_score : 0.0;
        c.is_end = last;
        nxt = (int)nodes.size();
        nodes.push_back(c);
        nodes[cur].next.emplace_back(tid, nxt);
      } else {
        BuildNode &x = nodes[nxt];
        x.token_score = std::max(sc, x.token_score);
        x.node_score = nodes[cur].node_score + x.token_score;
        if (last) {
          x.is_end = true;
          x.output_score = x.node_score;
        } else if (x.is_end) {
          x.output_score = x.node_score;
        }
      }
      cur = nxt;
    }
    ++n_phrases;
  }
  // failure + output links, breadth first (:91-137)
  std::deque<int> q;
  std::vector<int> bfs{0};
  for (const auto &e : nodes[0].next) {
    nodes[e.second].fail = 0;
    q.push_back(e.second);
  }
  while (!q.empty()) {
    const int cur = q.front();
    q.pop_front();
    bfs.push_back(cur);
    for (const auto &e : nodes[cur].next) {
      const int tid = e.first, ch = e.second;
      int f = nodes[cur].fail;
      int hit = nodes[f].find(tid);
      if (hit >= 0) {
        f = hit;
      } else {
        f = nodes[f].fail;
        hit = nodes[f].find(tid);
        while (hit < 0) {
          f = nodes[f].fail;
          if (nodes[f].token == -1) { hit = nodes[f].find(tid); break; }
          hit = nodes[f].find(tid);
        }
        if (hit >= 0) f = hit;
      }
      nodes[ch].fail = f;
      int out = f;
      while (!nodes[out].is_end) {
        out = nodes[out].fail;
        if (nodes[out].token == -1) { out = -1; break; }
      }
      nodes[ch].output = out;
      if (out >= 0) nodes[ch].output_score += nodes[out].output_score;
      q.push_back(ch);
    }
  }
  // flatten in BFS order, edges sorted by token id
  const int N = (int)nodes.size();
  std::vector<int> newid(N, -1);
  for (int i = 0; i < (int)bfs.size(); ++i) newid[bfs[i]] = i;
  edge_start.assign(N + 1, 0);
  edge_token.clear(); edge_child.clear();
  fail.assign(N, 0); token.assign(N, -1); is_end.assign(N, 0); output.assign(N, -1);
  token_score.assign(N, 0.0); node_score.assign(N, 0.0); output_score.assign(N, 0.0);
  for (int i = 0; i < N; ++i) {
    const BuildNode &nd = nodes[bfs[i]];
    edge_start[i] = (int)edge_token.size();
    std::vector<std::pair<int, int>> es = nd.next;
    std::sort(es.begin(), es.end());
    for (const auto &e : es) {
      edge_token.push_back(e.first);
      edge_child.push_back(newid[e.second]);
    }
    fail[i] = newid[nd.fail];
    token[i] = nd.token;
    is_end[i] = nd.is_end ? 1 : 0;
    output[i] = nd.output >= 0 ? newid[nd.output] : -1;
    token_score[i] = nd.token_score;
    node_score[i] = nd.node_score;
    output_score[i] = nd.output_score;
  }
  edge_start[N] = (int)edge_token.size();
}

}  // namespace b200asr

// ---- stand-alone host handle on the automaton (no recognizer, no CUDA): the same build and the same inline step function
// the search kernel runs, so the product's hotword graph can be held to the reference on a machine without a GPU
#include "../../include/b200asr.h"

extern "C" {

void *B200AsrHotwordGraphCreate(const int32_t *tokens, const int32_t *offsets, const float *scores, int32_t n_phrases) {
  if (n_phrases < 0 || (n_phrases > 0 && (!tokens || !offsets || !scores))) return nullptr;
  try {
    auto *g = new b200asr::ContextGraphHost();
    g->build(tokens, offsets, scores, n_phrases);
    return g;
  } catch (...) {
    return nullptr;
  }
}
void B200AsrHotwordGraphDestroy(void *g) { delete static_cast<b200asr::ContextGraphHost *>(g); }
int32_t B200AsrHotwordGraphNumNodes(const void *g) { return g ? static_cast<const b200asr::ContextGraphHost *>(g)->n_nodes() : -1; }
double B200AsrHotwordGraphStep(const void *g, int32_t state, int32_t token, int32_t *next_state) {
  const auto *h = static_cast<const b200asr::ContextGraphHost *>(g);
  if (!h || state < 0 || state >= h->n_nodes()) { if (next_state) *next_state = 0; return 0.0; }
  int nxt = 0;
  const double d = b200asr::cg_forward_one_step(h->view(), state, token, &nxt);
  if (next_state) *next_state = nxt;
  return d;
}
double B200AsrHotwordGraphFinalize(const void *g, int32_t state) {
  const auto *h = static_cast<const b200asr::ContextGraphHost *>(g);
  if (!h || state < 0 || state >= h->n_nodes()) return 0.0;
  return b200asr::cg_finalize(h->view(), state);
}

}  // extern "C"
