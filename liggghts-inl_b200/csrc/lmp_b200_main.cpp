// lmp_b200 -- command-line driver of the B200 DEM engine: `lmp_b200 -in in.deck [-device N] [-echo]`, the analogue of the
// reference's `lmp_<machine> -in in.deck` (src/main.cpp:40-70, src/lammps.cpp:120-330 for the switches).  Reads the deck
// through the input-script front end (dem_deck_file), so every `run` in the deck executes on the GPU, and closes with the
// reference's loop summary line (src/finish.cpp:100-130) for scripts that scrape it.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../include/dem_b200.h"

int main(int argc, char **argv)
{
  const char *in = nullptr; int device = 0; bool echo = false;
  for (int k = 1; k < argc; k++) {
    if ((!strcmp(argv[k], "-in") || !strcmp(argv[k], "-i")) && k + 1 < argc) in = argv[++k];
    else if (!strcmp(argv[k], "-device") && k + 1 < argc) device = atoi(argv[++k]);
    else if (!strcmp(argv[k], "-echo") || !strcmp(argv[k], "-e")) { echo = true; if (k + 1 < argc && argv[k + 1][0] != '-') k++; }
    else if (!strcmp(argv[k], "-suffix") || !strcmp(argv[k], "-sf")) { if (k + 1 < argc) k++; }  // `-suffix b200` of the reference binary: accepted
    else { fprintf(stderr, "usage: %s -in <input script> [-device N] [-echo]\n", argv[0]); return 2; }
  }
  if (!in) { fprintf(stderr, "usage: %s -in <input script> [-device N] [-echo]\n", argv[0]); return 2; }
  dem_engine *e = nullptr;
  if (dem_create(&e, device, 0, 1, nullptr, nullptr) != DEM_OK || !e) {
    fprintf(stderr, "ERROR: %s\n", e ? dem_last_error(e) : "engine creation failed (no usable sm_100 GPU?)");
    return 1;
  }
  dem_deck *d = nullptr;
  dem_deck_open(&d, e);
  dem_deck_screen(d, 1);  // thermo lines as they are produced, like the reference's screen
  printf("%s\n", dem_version());
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = dem_deck_file(d, in);
  if (rc != DEM_OK) { fprintf(stderr, "ERROR: %s\n", dem_deck_last_error(d)); dem_deck_close(d); dem_destroy(e); return 1; }  // error.cpp:160-186: message, exit(1)
  dem_stats st;
  dem_get_stats(e, &st);  // synchronises
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (echo && dem_deck_warnings(d)[0]) printf("%s", dem_deck_warnings(d));
  printf("Loop time of %g on 1 procs for %ld steps with %ld atoms\n", secs, dem_deck_ntimestep(d), st.nlocal);
  printf("Neighbor list builds = %ld\n", st.nbuilds);
  dem_deck_close(d); dem_destroy(e);
  return 0;
}
