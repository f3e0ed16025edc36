// gw-b200: Graph / DirectedGraph / UndirectedGraph, the graph containers Batch::get_graphs() fills and the Cython shim binds
// (pygenomeworks/genomeworks/cudapoa/graph.pxd:32-45). Same public surface as the reference's
// common/base/include/claraparabricks/genomeworks/utils/graph.hpp:50-110 (Graph: typedefs, get_adjacent_nodes, get_node_ids,
// get_edges, node labels), :226-283 (DirectedGraph: add_edge, dot / GFA serialisation) and :285-330 (UndirectedGraph);
// implemented from scratch on ordered containers (iteration order is therefore deterministic: by node id / by edge).
#pragma once

#include <algorithm>
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

/// Common part of the directed and the undirected graph: nodes are integers, edges carry an integer weight, nodes may
/// carry a text label.
class Graph
{
public:
    using node_id_t     = int32_t;
    using edge_weight_t = int32_t;
    using edge_t        = std::pair<node_id_t, node_id_t>;

    /// Nodes reachable over one edge from `node` (for the undirected graph: its neighbours), in insertion order.
    const std::vector<node_id_t>& get_adjacent_nodes(node_id_t node) const
    {
        static const std::vector<node_id_t> none;
        const auto it = adjacent_.find(node);
        return it == adjacent_.end() ? none : it->second;
    }

    /// Every node that is an end point of an edge or carries a label.
    const std::vector<node_id_t> get_node_ids() const
    {
        std::vector<node_id_t> ids;
        ids.reserve(adjacent_.size());
        for (const auto& kv : adjacent_)
            ids.push_back(kv.first);
        return ids;
    }

    /// All edges with their weights.
    const std::vector<std::pair<edge_t, edge_weight_t>> get_edges() const { return {edges_.begin(), edges_.end()}; }

    /// Sets the label of a node; the first label of a node stays (map insert semantics of the reference).
    void set_node_label(node_id_t node, const std::string& label)
    {
        labels_.emplace(node, label);
        adjacent_[node];
    }

    /// Label of a node, empty if it has none.
    std::string get_node_label(node_id_t node) const
    {
        const auto it = labels_.find(node);
        return it == labels_.end() ? std::string() : it->second;
    }

    virtual ~Graph() = default;

protected:
    Graph()             = default;
    Graph(const Graph&) = default;
    Graph(Graph&&)      = default;
    Graph& operator=(const Graph&) = default;
    Graph& operator=(Graph&&) = default;

    bool edge_exists(const edge_t& e) const { return edges_.find(e) != edges_.end(); }

    void link(node_id_t from, node_id_t to)
    {
        adjacent_[from].push_back(to);
        adjacent_[to]; // the far end is a known node as well
    }

    void labels_to_dot(std::ostringstream& os) const
    {
        for (const auto& kv : labels_)
            os << kv.first << " [label=\"" << kv.second << "\"];\n";
    }

    void edges_to_dot(std::ostringstream& os, const char* arrow) const
    {
        for (const auto& kv : edges_)
            os << kv.first.first << " " << arrow << " " << kv.first.second << " [label=\"" << kv.second << "\"];\n";
    }

    std::map<node_id_t, std::vector<node_id_t>> adjacent_;
    std::map<edge_t, edge_weight_t> edges_;
    std::map<node_id_t, std::string> labels_;
};

class DirectedGraph : public Graph
{
public:
    using Graph::edge_t;
    using Graph::edge_weight_t;
    using Graph::node_id_t;

    /// Adds edge from -> to with a weight (first insertion wins, as the reference's map insert does).
    void add_edge(node_id_t node_id_from, node_id_t node_id_to, edge_weight_t weight = 0)
    {
        const edge_t e(node_id_from, node_id_to);
        if (!edge_exists(e))
        {
            edges_.emplace(e, weight);
            link(node_id_from, node_id_to);
        }
    }

    /// Graphviz description: one line per labelled node, one per edge with its weight.
    std::string serialize_to_dot() const
    {
        std::ostringstream os;
        os << "digraph g {\n";
        labels_to_dot(os);
        edges_to_dot(os, "->");
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
};

class UndirectedGraph : public Graph
{
public:
    using Graph::edge_t;
    using Graph::edge_weight_t;
    using Graph::node_id_t;

    /// Adds the edge {from, to}; an edge that exists in either orientation is kept as it is.
    void add_edge(node_id_t node_id_from, node_id_t node_id_to, edge_weight_t weight = 0)
    {
        const edge_t e(node_id_from, node_id_to), r(node_id_to, node_id_from);
        if (!edge_exists(e) && !edge_exists(r))
        {
            edges_.emplace(e, weight);
            link(node_id_from, node_id_to);
            adjacent_[node_id_to].push_back(node_id_from);
        }
    }

    std::string serialize_to_dot() const
    {
        std::ostringstream os;
        os << "graph g {\n";
        labels_to_dot(os);
        edges_to_dot(os, "--");
        os << "}\n";
        return os.str();
    }
};

} // namespace genomeworks
} // namespace claraparabricks
