/*
 * oracle/cvshim/msvc_compat.hpp -- TEST INFRASTRUCTURE ONLY. Force-included (-include) in front of the reference's
 * OpenCVModified.cpp so that MSVC-only leniencies of its headers compile under GCC without editing them:
 *   - <algorithm> is reached transitively under MSVC (arcana/utils/algorithm.h uses std::nth_element without including it);
 *   - Image/ImageData.h:40 names `imageWidth` / `imageHeight` inside a constructor template that is never instantiated
 *     (MSVC does not look non-dependent names up until instantiation); two unused constants make the name lookup succeed.
 */
#pragma once
#include <algorithm>
#include <cstddef>
static const std::size_t imageWidth = 0, imageHeight = 0;
