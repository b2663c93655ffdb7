// The two boost::algorithm names InputDataPoroel.h:17 uses (split, is_any_of) — NOT boost; part of the deal.II API shim.
#pragma once
#include <string>
#include <vector>
namespace boost {
struct shim_any_of { std::string chars; };
inline shim_any_of is_any_of(const std::string& chars) { return shim_any_of{chars}; }
// token_compress_off semantics: adjacent delimiters give empty tokens, an empty input gives one empty token
template <class Container>
inline Container& split(Container& out, const std::string& input, const shim_any_of& pred) {
  out.clear();
  std::string cur;
  for (char ch : input) {
    if (pred.chars.find(ch) != std::string::npos) { out.push_back(cur); cur.clear(); }
    else cur.push_back(ch);
  }
  out.push_back(cur);
  return out;
}
}  // namespace boost
