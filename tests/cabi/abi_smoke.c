/* Pure-C consumer of include/tbk.h: proves that the header is valid C99 (no C++-isms, no torch or CUDA
 * types in the signatures), that a C program links against libtbk_b200.so, and that the argument checks of
 * the entry points answer without a GPU.  Built and run by tests/test_abi.py; no compute call is made. */
#include <stdio.h>
#include <string.h>
#include "tbk.h"

int main(void) {
  int bad = 0;
  if (tbk_version() < 100) { printf("version %d\n", tbk_version()); bad++; }
  if (tbk_last_error() == NULL) bad++;
  /* workspace queries are pure host arithmetic */
  if (tbk_eigh_workspace(200, 10, 1) == 0) { printf("no workspace for n = 200?\n"); bad++; }
  if (tbk_solve_workspace(2, 1024, 1) == 0) { printf("no workspace for the mesh kernel?\n"); bad++; }
  /* NULL handles / descriptors are rejected before anything touches the device */
  {
    tbk_model* m = NULL;
    if (tbk_model_create(NULL, &m) != TBK_ERR_ARG) { printf("model_create(NULL) accepted\n"); bad++; }
    if (strlen(tbk_last_error()) == 0) { printf("no error text\n"); bad++; }
    if (tbk_model_destroy(NULL) != TBK_OK) bad++;
  }
  {
    double k = 0.0, ev = 0.0;
    if (tbk_solve_k(NULL, &k, 1, &ev, 1, 1, NULL, 0, 0, NULL, 0, NULL) != TBK_ERR_ARG) { printf("solve_k(NULL) accepted\n"); bad++; }
    if (tbk_gen_ham(NULL, &k, 1, &ev, NULL) != TBK_ERR_ARG) { printf("gen_ham(NULL) accepted\n"); bad++; }
  }
  printf("abi_smoke: %s (tbk_version %d)\n", bad ? "FAILED" : "ok", tbk_version());
  return bad;
}
