"""Native host-side runtime pieces (C++, no CUDA dependency)."""
