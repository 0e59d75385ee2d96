/* placeholder translation unit for the gas-kinetic (cfg5) oracle; filled in later */
int fro_gks_available(void) { return 0; }
