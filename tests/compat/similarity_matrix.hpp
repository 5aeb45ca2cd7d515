// Stand-in for the reference's similarity_matrix.hpp: the declaration of the entry point.
#pragma once

#include "sequenced_data.hpp"
#include "util/mat.hpp"

#include <cstdint>
#include <string>
#include <vector>

Matd computeSimilarityMatrix(const std::vector<std::vector<PosData>> &pos_data, uint32_t num_cells,
                             uint32_t max_fragment_length, const std::vector<uint32_t> &group_id_to_pos,
                             double mutation_rate, double homozygous_rate, double seq_error_rate,
                             const uint32_t num_threads, const std::string &marker, const std::string &normalization);
