// goldpolish-index (B200 tree): drop-in for bcgsc/goldpolish src/goldpolish_index.cpp:3-17.
// Host only (no GPU work): builds and saves the byte-offset index the BF server loads.
#include "gp_host.hpp"

int main(int argc, char** argv)
{
  if (argc != 3) gph::die("Wrong args.");
  gph::SeqIndex::build(argv[1]).save(argv[2]);
  return 0;
}
