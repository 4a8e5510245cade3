// tokenizer.inl -- SURVEY.md section 8f rank 4: native restatement of the reference's
// host-side tokenizer (tokenizer.bin parse llama2.ts:441-449, bpe_encode :305-344, piece
// lookup for :501-503), for a node-free host.  Pure host code, no device work.
//
// Semantics kept: vocab strings are the UTF-8 DECODED token bytes (TextDecoder: malformed
// sequences become U+FFFD), `vocab.indexOf` returns the FIRST match, the text is walked by
// characters, the best-scoring adjacent pair is merged until none is in the vocab.  The
// reference's indexOf makes it O(n^2 * V); a hash map of first occurrences makes it O(n^2).
#include <string>
#include <unordered_map>
#include <vector>

struct l2b_tokenizer {
  std::vector<std::string> vocab;   // decoded (normalised) strings
  std::vector<float> scores;
  std::unordered_map<std::string, int> first;
  std::string err;
};

namespace {

// WHATWG UTF-8 decode -> re-encode: malformed input becomes U+FFFD (EF BF BD) per maximal subpart
std::string utf8_normalise(const unsigned char* s, size_t n) {
  std::string out;
  size_t i = 0;
  while (i < n) {
    const unsigned char c = s[i];
    int need = 0;
    unsigned lo = 0x80, hi = 0xBF;
    if (c < 0x80) { out.push_back((char)c); ++i; continue; }
    else if (c >= 0xC2 && c <= 0xDF) need = 1;
    else if (c >= 0xE0 && c <= 0xEF) { need = 2; if (c == 0xE0) lo = 0xA0; if (c == 0xED) hi = 0x9F; }
    else if (c >= 0xF0 && c <= 0xF4) { need = 3; if (c == 0xF0) lo = 0x90; if (c == 0xF4) hi = 0x8F; }
    else { out += "\xEF\xBF\xBD"; ++i; continue; }
    size_t j = i + 1;
    int got = 0;
    while (got < need && j < n) {
      const unsigned char d = s[j];
      const unsigned l = got == 0 ? lo : 0x80, h = got == 0 ? hi : 0xBF;
      if (d < l || d > h) break;
      ++j; ++got;
    }
    if (got == need) out.append((const char*)s + i, j - i);
    else out += "\xEF\xBF\xBD";
    i = j;
  }
  return out;
}

}  // namespace

L2B_API int l2b_tok_load(const uint8_t* data, uint64_t nbytes, int32_t vocab_size, l2b_tokenizer** out) {
  if (!data || !out || vocab_size <= 0) return L2B_EINVAL;
  *out = nullptr;
  l2b_tokenizer* t = new l2b_tokenizer();
  uint64_t p = 4;  // int32 max_token_length, ignored (llama2.ts:445)
  for (int i = 0; i < vocab_size; ++i) {
    if (p + 8 > nbytes) { delete t; return L2B_EINVAL; }
    float score; int32_t len;
    memcpy(&score, data + p, 4);
    memcpy(&len, data + p + 4, 4);
    p += 8;
    if (len < 0 || p + (uint64_t)len > nbytes) { delete t; return L2B_EINVAL; }
    t->scores.push_back(score);
    t->vocab.push_back(utf8_normalise(data + p, (size_t)len));
    t->first.emplace(t->vocab.back(), i);  // keeps the first occurrence, like indexOf
    p += (uint64_t)len;
  }
  *out = t;
  return L2B_OK;
}

// Returns the number of tokens, L2B_EINVAL when a character is not in the vocab (the
// reference throws, llama2.ts:310) or `cap` is too small.
L2B_API int l2b_tok_encode(const l2b_tokenizer* t, const char* text, int32_t* tokens, int32_t cap) {
  if (!t || !text || !tokens) return L2B_EINVAL;
  const std::string s = utf8_normalise((const unsigned char*)text, strlen(text));
  std::vector<int> tok;
  for (size_t i = 0; i < s.size();) {  // one lookup per character (JS: per UTF-16 code unit)
    const unsigned char c = (unsigned char)s[i];
    const size_t len = c < 0x80 ? 1 : c < 0xE0 ? 2 : c < 0xF0 ? 3 : 4;
    if (len == 4) return L2B_EINVAL;  // astral plane: JS looks up lone surrogates, never in the vocab
    auto it = t->first.find(s.substr(i, len));
    if (it == t->first.end()) return L2B_EINVAL;
    tok.push_back(it->second);
    i += len;
  }
  for (;;) {
    double best_score = -1e10;
    int best_id = -1, best_idx = -1;
    for (size_t i = 0; i + 1 < tok.size(); ++i) {
      auto it = t->first.find(t->vocab[tok[i]] + t->vocab[tok[i + 1]]);
      if (it != t->first.end() && (double)t->scores[it->second] > best_score) {
        best_score = (double)t->scores[it->second];
        best_id = it->second;
        best_idx = (int)i;
      }
    }
    if (best_idx == -1) break;
    tok[best_idx] = best_id;
    tok.erase(tok.begin() + best_idx + 1);
  }
  if ((int32_t)tok.size() > cap) return L2B_EINVAL;
  for (size_t i = 0; i < tok.size(); ++i) tokens[i] = tok[i];
  return (int)tok.size();
}

L2B_API const char* l2b_tok_piece(const l2b_tokenizer* t, int32_t id) {
  if (!t || id < 0 || id >= (int32_t)t->vocab.size()) return nullptr;
  return t->vocab[id].c_str();
}

// pieces may contain NUL bytes (the raw-byte tokens 3..258): length in bytes
L2B_API int32_t l2b_tok_piece_len(const l2b_tokenizer* t, int32_t id) {
  if (!t || id < 0 || id >= (int32_t)t->vocab.size()) return -1;
  return (int32_t)t->vocab[id].size();
}

L2B_API float l2b_tok_score(const l2b_tokenizer* t, int32_t id) {
  if (!t || id < 0 || id >= (int32_t)t->scores.size()) return 0.f;
  return t->scores[id];
}

L2B_API void l2b_tok_free(l2b_tokenizer* t) { delete t; }
