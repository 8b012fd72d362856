#pragma once
#include "cereal/cereal.hpp"
class IndexHeader {
  public:
    bool bigSA() const { return bigSA_; }
    template <typename Archive> void serialize(Archive& ar) { ar(cereal::make_nvp("bigSA", bigSA_)); }
  private:
    bool bigSA_{false};
};
