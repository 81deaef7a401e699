// Library-level entry points: version, last-error string.
#include <stdio.h>

#include "common.cuh"

static thread_local char g_last_error[512] = "";

void mmemo_set_error(const char* what, const char* file, int line) {
  snprintf(g_last_error, sizeof(g_last_error), "%s (%s:%d)", what ? what : "?", file, line);
}

int g_mm_pdl = 1;

extern "C" {
int mmemo_version(void) { return 100; }
int mmemo_set_pdl(int enabled) {
  g_mm_pdl = enabled ? 1 : 0;
  return MMEMO_OK;
}
const char* mmemo_last_error(void) { return g_last_error; }
}
