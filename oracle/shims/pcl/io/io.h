// shim: forwards to the inert third-party stand-ins (oracle/_ref build only)
#include "../../gndt_shim_core.h"
