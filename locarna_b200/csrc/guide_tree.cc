// Guide tree of mlocarna's all-vs-all stage: UPGMA over the pairwise score matrix, with the reference's tie rules.
// Reference: lib/perl/MLocarna/Tree.pm:181-262 (_upgma_dist), :265-295 (_scores_to_dists), :97-146 (to_newick, label quoting),
// caller src/Utils/mlocarna:2353-2386 (diagonal 0, "result.tree" = newick + ";").
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/locarna_b200.h"

namespace {

struct Node { int left = -1, right = -1; std::string label, length; bool has_length = false; };

std::string quote_label(const std::string &label) {  // Tree.pm:131-146
    bool ticks = label.find('\'') != std::string::npos || label.find(':') != std::string::npos;
    std::string out;
    for (char ch : label) { out += ch; if (ch == '\'') out += '\''; }
    return ticks ? "'" + out + "'" : out;
}

void to_newick(const std::vector<Node> &nodes, int k, std::string &out) {  // Tree.pm:97-119
    const Node &nd = nodes[k];
    if (nd.left >= 0) {
        out += '(';
        to_newick(nodes, nd.left, out);
        out += ',';
        to_newick(nodes, nd.right, out);
        out += ')';
    }
    out += quote_label(nd.label);
    if (nd.has_length) out += ":" + nd.length;
}

}  // namespace

extern "C" int lb200_upgma_newick(int n, const char *const *names, const int64_t *scores, char *out, size_t out_cap) {
    if (n < 1 || !names || !scores || !out) return LB200_ERR_ARG;
    // scores -> distances: max over the upper triangle (initialised with [0][0]) minus score (Tree.pm:265-295)
    std::vector<std::vector<double>> dist(n, std::vector<double>(n));
    double mx = (double)scores[0];
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) if ((double)scores[(size_t)i * n + j] >= mx) mx = (double)scores[(size_t)i * n + j];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) dist[i][j] = mx - (double)scores[(size_t)i * n + j];
    std::vector<int> clusters(n), sizes(n, 1), tree_of(n);
    std::vector<double> heights(n, 0.0);
    std::vector<Node> nodes(n);
    for (int i = 0; i < n; i++) { clusters[i] = i; tree_of[i] = i; nodes[i].label = names[i]; }  // quoted when printed (the reference quotes, re-parses, and quotes again on output)
    const double INF = 1e10;
    int index = 1;
    int last = n - 1;  // $#clusters
    while (last > 0) {
        int min_i = -1, min_j = -1;
        double min_dist = INF;
        for (int i = 0; i <= last; i++)
            for (int j = i + 1; j <= last; j++) {
                const double d = dist[clusters[i]][clusters[j]];
                if (d < min_dist) { min_i = i; min_j = j; min_dist = d; }  // first strict minimum in list order
            }
        if (min_i < 0) return LB200_ERR_ARG;  // all distances >= 1e10
        const int ci = clusters[min_i], cj = clusters[min_j];
        clusters[min_j] = clusters[last];
        clusters[min_i] = clusters[0];
        clusters[0] = ci;
        for (int i = 1; i < last; i++) {
            const double v = (sizes[ci] * dist[ci][clusters[i]] + sizes[cj] * dist[cj][clusters[i]]) / (sizes[ci] + sizes[cj]);
            dist[clusters[0]][clusters[i]] = v;
            dist[clusters[i]][clusters[0]] = v;
        }
        const double height = min_dist / 2.0;
        char buf[64];
        snprintf(buf, sizeof buf, "%.3f", height - heights[ci]);
        nodes[tree_of[ci]].length = buf; nodes[tree_of[ci]].has_length = true;
        snprintf(buf, sizeof buf, "%.3f", height - heights[cj]);
        nodes[tree_of[cj]].length = buf; nodes[tree_of[cj]].has_length = true;
        Node parent;
        parent.left = tree_of[ci]; parent.right = tree_of[cj];
        parent.label = std::to_string(index++);
        nodes.push_back(parent);
        tree_of[clusters[0]] = (int)nodes.size() - 1;
        sizes[clusters[0]] = sizes[ci] + sizes[cj];
        heights[clusters[0]] = height;
        last--;
    }
    std::string nw;
    to_newick(nodes, tree_of[clusters[0]], nw);
    if (nw.size() + 1 > out_cap) return LB200_ERR_ARG;
    memcpy(out, nw.c_str(), nw.size() + 1);
    return LB200_OK;
}
