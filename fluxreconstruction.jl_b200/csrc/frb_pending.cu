// (all entry points of include/frb200.h are implemented; this translation unit is intentionally empty)
