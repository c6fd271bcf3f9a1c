// Host build of chess_b200/csrc/parse_number.cuh for the CPU test suite: reads one field per line
// on stdin, prints "<rc> <bits of the double as hex>" per line (mode "f") or "<rc> <int>" (mode "i").
#include <cstdio>
#include <cstring>
#include <string>
#include <iostream>

#include "../chess_b200/csrc/parse_number.cuh"

static const uint64_t kPow5[] = CHESS_POW5_TABLE;
static const short kExp5[] = CHESS_POW5_EXP;
static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

int main(int argc, char **argv) {
  const bool ints = argc > 1 && std::strcmp(argv[1], "i") == 0;
  std::string line;
  while (std::getline(std::cin, line)) {
    if (ints) {
      int32_t v = 0;
      const int rc = chess_parse_index(line.data(), (int)line.size(), &v);
      std::printf("%d %d\n", rc, rc == PARSE_OK ? v : 0);
    } else {
      double d = 0.0;
      const int rc = chess_parse_double(line.data(), (int)line.size(), kPow5, kExp5, kPow10, &d);
      uint64_t bits = 0;
      std::memcpy(&bits, &d, 8);
      std::printf("%d %016llx\n", rc, rc == PARSE_OK ? (unsigned long long)bits : 0ull);
    }
  }
  return 0;
}
