/* minimal.c -- the engine from plain C: the calls a RAPID / FLANN user makes, through include/sffg.h.
 *
 *   gcc -std=c99 -I include examples/minimal.c -L space_filling_forest_star_b200 -l:libsffg.so \
 *       -Wl,-rpath,$PWD/space_filling_forest_star_b200 -o minimal && ./minimal
 *
 * One obstacle triangle, a one-triangle robot, three poses, one edge, a four-node index.  Exit code 0 when every
 * answer is the expected one, 2 when no sm_100 GPU is present (the library has no CPU path), 1 on any other error. */
#include <stdio.h>

#include "sffg.h"

static int fail(const char *what) {
  fprintf(stderr, "%s: %s\n", what, sffg_last_error());
  return 1;
}

int main(void) {
  const double obstacle[9] = {0, 0, 0, 4, 0, 0, 0, 4, 0};             /* RAPID_model::AddTri(p1, p2, p3) of the map  */
  const double robot[9] = {-0.5, 0, -0.5, 0.5, 0, -0.5, 0, 0, 0.5};   /* ... and of the robot                        */
  const double poses[3][6] = {{1, 1, 0, 0, 0, 0},                     /* x y z yaw pitch roll: cuts the map triangle  */
                              {1, 1, 3, 0, 0, 0},                     /* above it                                     */
                              {1, 1, 0.4, 0, 0, 0}};                  /* still reaches down through it                */
  const double start[6] = {1, 1, 3, 0, 0, 0}, end[6] = {1, 1, -3, 0, 0, 0};
  const float nodes[4][6] = {{0, 0, 0, 0, 0, 0}, {1, 0, 0, 0, 0, 0}, {0, 2, 0, 0, 0, 0}, {0, 0, 3, 3.0f, 0, 0}};
  const float query[6] = {0.9f, 0.1f, 0, 0, 0, 0};
  sffg_env *env = 0;
  sffg_index *idx = 0;
  uint8_t verdict[3], is_free = 9;
  int32_t first_hit = -1, ids[2];
  float d2[2];
  int rc = sffg_init(-1);
  if (rc == SFFG_ERR_NO_DEVICE) {
    fprintf(stderr, "%s\n", sffg_last_error());
    return 2;
  }
  if (rc != SFFG_OK) return fail("sffg_init");
  if (sffg_env_create(obstacle, 1, robot, 1, &env) != SFFG_OK) return fail("sffg_env_create");
  if (sffg_collide_poses_f64(env, &poses[0][0], 3, verdict) != SFFG_OK) return fail("sffg_collide_poses_f64");
  if (sffg_check_edges(env, start, end, 1, 0.1, SFFG_ROT_REFERENCE, &is_free, &first_hit) != SFFG_OK) return fail("sffg_check_edges");
  if (sffg_index_create(6, &idx) != SFFG_OK) return fail("sffg_index_create");
  if (sffg_index_add(idx, &nodes[0][0], 4) != SFFG_OK) return fail("sffg_index_add");
  if (sffg_knn(idx, query, 1, 2, ids, d2) != SFFG_OK) return fail("sffg_knn");
  printf("verdicts %d %d %d, edge free %d (first colliding sample %d), nearest nodes %d %d (d2 %.3f %.3f)\n", verdict[0], verdict[1],
         verdict[2], is_free, first_hit, ids[0], ids[1], d2[0], d2[1]);
  sffg_index_destroy(idx);
  sffg_env_destroy(env);
  /* the robot triangle spans z in [-0.5, 0.5] about its origin: it cuts the map at z = 0 and 0.4, not at 3; the edge
   * runs from z = 3 down to -3 in steps of 0.1 and first touches the map when the robot's top reaches it (sample 25,
   * z = 0.5); node 1 (distance^2 0.02) and node 0 (0.82) are the nearest two */
  return (verdict[0] == 1 && verdict[1] == 0 && verdict[2] == 1 && is_free == 0 && first_hit == 25 && ids[0] == 1 && ids[1] == 0) ? 0 : 1;
}
