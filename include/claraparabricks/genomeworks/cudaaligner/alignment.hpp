// gw-b200: cudaaligner::Alignment -- the result object of the reference API
// (cudaaligner/include/claraparabricks/genomeworks/cudaaligner/alignment.hpp:37-111). Results of the banded Myers engine are
// run-length encoded (actions + run lengths); CIGAR conversion follows cudaaligner/src/alignment_impl.cpp:99-153.
#pragma once

#include "cudaaligner.hpp"

#include <ostream>
#include <string>
#include <vector>

namespace claraparabricks
{
namespace genomeworks
{
namespace cudaaligner
{

typedef struct FormattedAlignment
{
    std::string query;
    std::string pairing;
    std::string target;
    uint32_t linebreak_after = 80;
} FormattedAlignment;

inline std::ostream& operator<<(std::ostream& os, const FormattedAlignment& f)
{
    const size_t n    = f.query.size();
    const size_t step = f.linebreak_after == 0 ? (n == 0 ? 1 : n) : f.linebreak_after;
    for (size_t i = 0; i < n; i += step)
        os << f.query.substr(i, step) << '\n' << f.pairing.substr(i, step) << '\n' << f.target.substr(i, step) << "\n\n";
    return os;
}

class Alignment
{
public:
    virtual ~Alignment()                                                              = default;
    virtual const std::string& get_query_sequence() const                            = 0;
    virtual const std::string& get_target_sequence() const                           = 0;
    virtual std::string convert_to_cigar(CigarFormat format = CigarFormat::basic) const = 0;
    virtual AlignmentType get_alignment_type() const                                 = 0;
    virtual bool is_optimal() const                                                  = 0;
    virtual StatusType get_status() const                                            = 0;
    virtual const std::vector<AlignmentState>& get_alignment() const                 = 0;
    virtual const std::vector<int8_t>& get_actions() const                           = 0;
    virtual const std::vector<int32_t>& get_runlengths() const                       = 0;
    virtual int32_t get_edit_distance() const                                        = 0;
    virtual FormattedAlignment format_alignment(int32_t maximal_line_length = 80) const = 0;
};

namespace detail
{
class AlignmentB200 : public Alignment
{
public:
    AlignmentB200(std::string q, std::string t)
        : query_(std::move(q))
        , target_(std::move(t))
    {
    }
    void set_type(AlignmentType t) { type_ = t; }
    void set(StatusType st, bool optimal, std::vector<int8_t> actions, std::vector<int32_t> runs)
    {
        status_     = st;
        is_optimal_ = optimal;
        actions_    = std::move(actions);
        runs_       = std::move(runs);
        type_       = AlignmentType::global_alignment;
    }
    /// AlignerGlobal results are one AlignmentState per step (aligner_global.cpp:162-190), not run-length encoded
    void expand()
    {
        expanded_.clear();
        for (size_t k = 0; k < actions_.size(); ++k)
            expanded_.insert(expanded_.end(), static_cast<size_t>(runs_[k]), static_cast<AlignmentState>(actions_[k]));
    }
    const std::string& get_query_sequence() const override { return query_; }
    const std::string& get_target_sequence() const override { return target_; }
    AlignmentType get_alignment_type() const override { return type_; }
    bool is_optimal() const override { return is_optimal_; }
    StatusType get_status() const override { return status_; }
    const std::vector<AlignmentState>& get_alignment() const override { return expanded_; } // empty for RLE results, as in the reference
    const std::vector<int8_t>& get_actions() const override { return actions_; }
    const std::vector<int32_t>& get_runlengths() const override { return runs_; }
    std::string convert_to_cigar(CigarFormat format = CigarFormat::basic) const override
    {
        std::string cigar;
        if (actions_.empty())
            return cigar;
        if (format == CigarFormat::extended)
        {
            static const char ext[4] = {'=', 'X', 'I', 'D'};
            for (size_t i = 0; i < actions_.size(); ++i)
                cigar += std::to_string(runs_[i]) + ext[actions_[i] & 3];
            return cigar;
        }
        static const char bas[4] = {'M', 'M', 'I', 'D'};
        char last                = bas[actions_[0] & 3];
        int64_t count            = runs_[0];
        for (size_t i = 1; i < actions_.size(); ++i)
        {
            const char c = bas[actions_[i] & 3];
            if (c == last)
            {
                count += runs_[i];
            }
            else
            {
                cigar += std::to_string(count) + last;
                last  = c;
                count = runs_[i];
            }
        }
        cigar += std::to_string(count) + last;
        return cigar;
    }
    int32_t get_edit_distance() const override
    {
        int32_t d = 0;
        for (size_t i = 0; i < actions_.size(); ++i)
            if (actions_[i] != static_cast<int8_t>(AlignmentState::match))
                d += runs_[i];
        return d;
    }
    FormattedAlignment format_alignment(int32_t maximal_line_length = 80) const override
    {
        FormattedAlignment f;
        f.linebreak_after = maximal_line_length < 0 ? 0u : static_cast<uint32_t>(maximal_line_length);
        size_t qi = 0, ti = 0;
        for (size_t k = 0; k < actions_.size(); ++k)
        {
            for (int32_t r = 0; r < runs_[k]; ++r)
            {
                switch (actions_[k])
                {
                case AlignmentState::match:
                case AlignmentState::mismatch:
                    f.query += query_[qi++];
                    f.target += target_[ti++];
                    f.pairing += actions_[k] == AlignmentState::match ? '|' : 'x';
                    break;
                case AlignmentState::insertion:
                    f.query += '-';
                    f.target += target_[ti++];
                    f.pairing += ' ';
                    break;
                default:
                    f.query += query_[qi++];
                    f.target += '-';
                    f.pairing += ' ';
                    break;
                }
            }
        }
        return f;
    }

private:
    std::string query_, target_;
    StatusType status_    = StatusType::uninitialized;
    AlignmentType type_   = AlignmentType::unset;
    bool is_optimal_      = false;
    std::vector<int8_t> actions_;
    std::vector<int32_t> runs_;
    std::vector<AlignmentState> expanded_;
};
} // namespace detail

} // namespace cudaaligner
} // namespace genomeworks
} // namespace claraparabricks
