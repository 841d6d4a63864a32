// ABI identification for include/ccd_b200.h
extern "C" int ccd_abi_version(void) { return 1; }
