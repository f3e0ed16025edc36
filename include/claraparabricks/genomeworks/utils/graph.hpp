// gw-b200: DirectedGraph, the graph container Batch::get_graphs() fills. Same public surface as the reference's
// common/base/include/claraparabricks/genomeworks/utils/graph.hpp:226-283 (node ids, weighted edges, node labels,
// dot / GFA serialisation); implemented from scratch on ordered containers.
#pragma once

#include <cstdint>
#include <map>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace claraparabricks
{
namespace genomeworks
{

class DirectedGraph
{
public:
    using node_id_t     = int32_t;
    using edge_weight_t = int32_t;
    using edge_t        = std::pair<node_id_t, node_id_t>;

    /// Adds edge src -> dst with a weight (first insertion wins, as the reference's map insert does).
    void add_edge(node_id_t src, node_id_t dst, edge_weight_t weight = 0)
    {
        const edge_t e(src, dst);
        if (edges_.find(e) == edges_.end())
        {
            edges_.emplace(e, weight);
            adjacent_[src].push_back(dst);
            adjacent_[dst]; // make sure the sink is a known node
        }
    }
    void set_node_label(node_id_t node, const std::string& label)
    {
        labels_.emplace(node, label);
        adjacent_[node];
    }
    std::string get_node_label(node_id_t node) const
    {
        auto it = labels_.find(node);
        return it == labels_.end() ? std::string() : it->second;
    }
    const std::vector<node_id_t>& get_adjacent_nodes(node_id_t node) const
    {
        static const std::vector<node_id_t> empty;
        auto it = adjacent_.find(node);
        return it == adjacent_.end() ? empty : it->second;
    }
    std::vector<node_id_t> get_node_ids() const
    {
        std::vector<node_id_t> ids;
        for (const auto& kv : adjacent_)
            ids.push_back(kv.first);
        return ids;
    }
    std::vector<std::pair<edge_t, edge_weight_t>> get_edges() const { return {edges_.begin(), edges_.end()}; }

    /// Graphviz description: one line per labelled node, one per edge with its weight.
    std::string serialize_to_dot() const
    {
        std::ostringstream os;
        os << "digraph g {\n";
        for (const auto& kv : labels_)
            os << kv.first << " [label=\"" << kv.second << "\"];\n";
        for (const auto& kv : edges_)
            os << kv.first.first << " -> " << kv.first.second << " [label=\"" << kv.second << "\"];\n";
        os << "}\n";
        return os.str();
    }
    /// GFA 1.0: S lines for nodes (label as sequence), L lines for edges.
    std::string serialize_to_gfa() const
    {
        std::ostringstream os;
        os << "H\tVN:Z:1.0\n";
        for (const auto& kv : adjacent_)
            os << "S\t" << kv.first << "\t" << get_node_label(kv.first) << "\n";
        for (const auto& kv : edges_)
            os << "L\t" << kv.first.first << "\t+\t" << kv.first.second << "\t+\t0M\tRC:i:" << kv.second << "\n";
        return os.str();
    }

private:
    std::map<node_id_t, std::vector<node_id_t>> adjacent_;
    std::map<edge_t, edge_weight_t> edges_;
    std::map<node_id_t, std::string> labels_;
};

} // namespace genomeworks
} // namespace claraparabricks
