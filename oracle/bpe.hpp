// TEST INFRASTRUCTURE (oracle): byte-level BPE token count, the CPU side of K4.  Only tests/, __graft_entry__.smoke() and
// bench.py's CPU legs may use it; the product never does.
//
// SELF-ORACLE, NOT REFERENCE PARITY: the reference has no tokenizer (token counts come from provider responses).
// This restates what HuggingFace `tokenizers` 0.22 computes for the tokenizer tests/golden/make_bpe_vocab.py defines
// (Split(" ", merged_with_next) + ByteLevel(use_regex=False) + BPE without dropout), and is pinned by the library's own answers
// for 614 texts (tests/golden/bpe_cases.json) and, where the library is importable, live on random texts.
//   pre-tokenisation: every space starts a new piece; a piece is its first byte plus the bytes up to the next space
//   model: inside a piece, merge the adjacent pair of lowest rank, leftmost first among equal ranks, until no pair is mergeable
#pragma once
#include <cstdint>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace oracle {

struct BpeVocab {
  uint16_t byte_to_id[256];
  std::unordered_map<uint32_t, uint32_t> merges;   // (a << 16 | b) → rank << 16 | merged id
};

inline void bpe_load(BpeVocab& V, const uint16_t* byte_to_id, const uint32_t* merges /* n × (a, b, merged) */, uint32_t n) {
  for (int i = 0; i < 256; i++) V.byte_to_id[i] = byte_to_id[i];
  V.merges.clear(); V.merges.reserve(n * 2);
  for (uint32_t r = 0; r < n; r++) V.merges.emplace((merges[3 * r] << 16) | merges[3 * r + 1], (r << 16) | merges[3 * r + 2]);
}

inline uint32_t bpe_count(const BpeVocab& V, std::string_view text) {
  uint32_t total = 0;
  std::vector<uint16_t> s;
  size_t i = 0; const size_t n = text.size();
  while (i < n) {
    size_t j = i + 1; while (j < n && text[j] != ' ') j++;
    s.clear(); for (size_t k = i; k < j; k++) s.push_back(V.byte_to_id[(unsigned char)text[k]]);
    for (;;) {
      uint32_t best = 0xffffffffu; size_t at = 0; uint32_t merged = 0;
      for (size_t k = 0; k + 1 < s.size(); k++) {
        auto it = V.merges.find(((uint32_t)s[k] << 16) | s[k + 1]);
        if (it != V.merges.end() && (it->second >> 16) < best) { best = it->second >> 16; at = k; merged = it->second & 0xffffu; }
      }
      if (best == 0xffffffffu) break;
      s[at] = (uint16_t)merged; s.erase(s.begin() + at + 1);
    }
    total += (uint32_t)s.size();
    i = j;
  }
  return total;
}

}  // namespace oracle
