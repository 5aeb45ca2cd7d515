// Stand-in for the reference's sequenced_data.hpp (API-compatible subset, written from the interface
// description in SURVEY.md §8a1) so that the drop-in shim can be built and tested where the SECEDO
// source tree is not available. When building inside SECEDO, put its own headers first on the
// include path instead of this directory.
#pragma once

#include <cstdint>
#include <utility>
#include <vector>

struct PosData {
    PosData(uint32_t pos, std::vector<uint32_t> ids, std::vector<uint16_t> gids_bases)
        : position(pos), read_ids(std::move(ids)), group_ids_bases(std::move(gids_bases)) {}

    uint32_t position;
    std::vector<uint32_t> read_ids;
    // 14-bit group id << 2 | 2-bit base
    std::vector<uint16_t> group_ids_bases;

    uint16_t group_id(uint32_t i) const { return group_ids_bases[i] >> 2; }
    uint8_t base(uint32_t i) const { return group_ids_bases[i] & 3; }
    uint32_t size() const { return static_cast<uint32_t>(read_ids.size()); }
    bool operator==(const PosData &o) const {
        return position == o.position && read_ids == o.read_ids && group_ids_bases == o.group_ids_bases;
    }
};
