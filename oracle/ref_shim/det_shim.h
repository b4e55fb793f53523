// TEST INFRASTRUCTURE ONLY (oracle build).  Forced include (-include) for the reference's
// tree-build translation unit.
//
// The reference sorts (key, id) pairs by key only with unstable sorts
// (deltapq_create_approx_tree.h:524-528, 1077-1079, 1423-1425), so its edge list depends
// on the OpenMP thread count.  Pulling in every std header first and then renaming
// `sort` makes std::sort -> std::stable_sort and __gnu_parallel::sort ->
// __gnu_parallel::stable_sort without touching the reference sources.  The resulting
// "canonical" tree (ties broken by ascending id / emission order) is thread-count
// independent and is the oracle for bit-exact edges / QNode / compressed-stream files.
#ifndef DPQ_ORACLE_DET_SHIM_H
#define DPQ_ORACLE_DET_SHIM_H
#include <opencv2/opencv.hpp>
#include <algorithm>
#include <parallel/algorithm>
#include <bitset>
#include <unordered_map>
#include <unordered_set>
#include <omp.h>
#define sort stable_sort
#endif
