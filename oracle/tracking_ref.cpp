/*
 * oracle/tracking_ref.cpp -- TEST INFRASTRUCTURE ONLY. C wrapper around the REFERENCE's own header
 * Core/MAGESLAM/Source/Map/MappingMath.h (dependency-free), compiled where it lies into oracle/_ref/libtracking_ref.so.
 * Pins the restated ComputeOctave of tracking_oracle.cpp and the DMin/DMax helpers the test scenes are generated with.
 */
#include <cstdint>     /* the header uses uint64_t without including it (MSVC pulls it in transitively) */
#include <Map/MappingMath.h>

extern "C" {
int   trkref_compute_octave(float distance, float dmin, float scale_factor) { return mage::ComputeOctave(distance, dmin, scale_factor); }
float trkref_compute_dmax(float distance, int octave, int max_octave, float scale_factor) { return mage::ComputeDMax(distance, octave, max_octave, scale_factor); }
float trkref_compute_dmin(float distance, int octave, float scale_factor) { return mage::ComputeDMin(distance, octave, scale_factor); }
}
