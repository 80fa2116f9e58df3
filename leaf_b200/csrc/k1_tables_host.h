// Host-side construction of the K1 lookup tables (static data from k1_tables.inc + the merge hash table).
// Shared by the engine (which uploads them to the device) and by the CPU test harness.
#pragma once
#include <stdint.h>
#include <vector>

#include "k1_core.cuh"

namespace leaf {
namespace k1host {
#include "k1_tables.inc"
}

constexpr uint32_t K1_MERGE_BITS = 17;   // 131072 slots * 8 B = 1 MiB, load factor 0.37 (L2 resident)

// merge_pairs[r] = (left << 16) | right ; returns the open-addressing table
inline std::vector<uint64_t> k1_build_merge_table(const uint32_t* merge_pairs, int n_merges) {
  std::vector<uint64_t> tab(1u << K1_MERGE_BITS, K1_SLOT_EMPTY);
  const uint32_t mask = (1u << K1_MERGE_BITS) - 1u;
  for (int r = 0; r < n_merges; ++r) {
    const uint32_t key = merge_pairs[r];
    uint32_t slot = k1_hash(key, K1_MERGE_BITS);
    while (tab[slot] != K1_SLOT_EMPTY) slot = (slot + 1u) & mask;
    tab[slot] = (static_cast<uint64_t>(key) << 32) | static_cast<uint64_t>(r);
  }
  return tab;
}

inline K1Tables k1_host_tables(const uint64_t* merge_tab) {
  K1Tables T;
  T.byte_id = k1host::k1_host_byte_id;
  T.cls = k1host::k1_host_class;
  T.ws = k1host::k1_host_ws;
  T.lower = k1host::k1_host_lower;
  T.numref = k1host::k1_host_numref;
  T.ent_off = k1host::k1_host_ent_off;
  T.ent_len = k1host::k1_host_ent_len;
  T.ent_val = k1host::k1_host_ent_val;
  T.ent_blob = k1host::k1_host_ent_blob;
  T.n_ent = K1_N_ENTITIES;
  T.merge_tab = merge_tab;
  T.merge_bits = K1_MERGE_BITS;
  return T;
}

}  // namespace leaf
