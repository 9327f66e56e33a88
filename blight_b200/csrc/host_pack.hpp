// host_pack.hpp — ASCII -> 2-bit packer of the host-buffer entry points (host_pack.cpp).
#pragma once
#include <cstdint>

namespace blight {

// Packs text[0, n_bases) into words[0, ceil(n_bases / 16)): base i -> bits 30 - 2 (i & 15) .. of word i >> 4, code
// (c >> 1) & 3 (nuc2int, kmer.h:56-69); bases past the end of the last word read as A. Returns false if any byte is not
// one of ACGTacgt (the words are then meaningless for that byte: send the block as ASCII instead).
bool pack2_block(const char* text, uint64_t n_bases, uint32_t* words);

}  // namespace blight
