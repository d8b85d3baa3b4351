// Error reporting shared by every translation unit of libnerfloam_b200.so: a thread-local message
// behind nl_last_error(); entry points return a negative status instead of exiting the process.
#pragma once
#include "../../include/nerfloam_b200.h"

int nl_set_error(const char *msg);                 // stores msg, returns NL_ERR_INVALID
int nl_set_error_code(int code, const char *msg);  // stores msg, returns code
