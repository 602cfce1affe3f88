// sc_hla_count — command-line front end (src/main.rs:135-147): same flags, same output layout.
#include <cstdio>

#include "../include/hlacount.h"

int main(int argc, char** argv) {
  int rc = hlac_main(argc, (const char* const*)argv);
  if (rc != 0) {
    printf("Failed with error: %s\n", hlac_last_error());
    return 1;
  }
  return 0;
}
