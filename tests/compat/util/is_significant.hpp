// Stand-in for the reference's util/is_significant.hpp: same public interface of Filter.
#pragma once

#include "sequenced_data.hpp"

#include <array>
#include <limits>
#include <string>
#include <vector>

constexpr uint16_t NO_POS = std::numeric_limits<uint16_t>::max() >> 2;

class Filter {
    std::vector<double> log_factorial;
    double theta;
    uint8_t cell_proportion;
    double log_theta_3;
    double log_one_minus_theta;

  public:
    Filter(double theta, uint8_t cell_proportion = 4);
    bool is_significant(std::array<uint16_t, 4> &base_count);
    bool is_significant(const PosData &pos_data, uint16_t *coverage);
    std::pair<std::vector<std::vector<PosData>>, double> filter(const std::vector<std::vector<PosData>> &pos_data,
                                                                const std::vector<uint32_t> &id_to_pos,
                                                                const std::string &marker, uint32_t num_threads);
    double log_fact(uint32_t n);
};
