// Test tooling: host build of aigw_b200/csrc/shortest_f64.cuh; reads decimal strings on stdin, prints 1 / 0 per line.
#include <cstdio>
#include <iostream>
#include <string>
#include "../aigw_b200/csrc/shortest_f64.cuh"
int main() {
  std::string l;
  while (std::getline(std::cin, l)) printf("%d\n", aigw::shortest_f64_decimal((const uint8_t*)l.data(), (uint32_t)l.size()) ? 1 : 0);
}
