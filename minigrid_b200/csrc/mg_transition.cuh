// mg_transition.cuh — MiniGridEnv.step's state transition (minigrid_env.py:525-588) for one environment,
// on the byte-coded cell in front of the agent. Pure register logic: the caller reads the front cell and
// writes it back when `newc != fc`.
#pragma once
#include "mg_common.cuh"

namespace mg {

struct StepOut {
  uint32_t newc;        // front cell after the action (== fc when unchanged)
  uint32_t terminated;  // 0/1
  uint32_t goal;        // reached the goal: reward = _reward() (minigrid_env.py:240-245)
  uint32_t bad_action;  // action outside 0..6 (ValueError, :584-585)
};

// DIR_TO_VEC (core/constants.py:49-58)
MG_HD void front_pos(const Geom &g, int ax, int ay, int dir, int &fx, int &fy) {
  const int dx = (dir == 0) - (dir == 2), dy = (dir == 1) - (dir == 3);
  fx = clampi(ax + dx, 0, g.W - 1);
  fy = clampi(ay + dy, 0, g.H - 1);
}

MG_HD StepOut transition(int action, uint32_t fc, int fx, int fy, int &ax, int &ay, int &dir, uint32_t &carry) {
  StepOut o;
  o.newc = fc; o.terminated = 0; o.goal = 0; o.bad_action = 0;
  const uint32_t t4 = fc & 15u, col = (fc >> 4) & 7u;
  if (action == A_LEFT) dir = (dir + 3) & 3;                 // :541-544
  else if (action == A_RIGHT) dir = (dir + 1) & 3;           // :547-548
  else if (action == A_FORWARD) {                            // :551-558
    // can_overlap: None, Goal, Floor, Lava, open Door (world_object.py:45,113,128,141,177)
    if ((0x031Au >> t4) & 1u) { ax = fx; ay = fy; }
    if (t4 == T_GOAL) { o.terminated = 1; o.goal = 1; }
    if (t4 == T_LAVA) o.terminated = 1;
  } else if (action == A_PICKUP) {                           // :561-566, can_pickup: Key, Ball, Box
    if (t4 >= T_KEY && t4 <= T_BOX && carry == 0) { carry = fc & 0x7Fu; o.newc = CODE_EMPTY; }
  } else if (action == A_DROP) {                             // :569-573
    if (t4 == T_EMPTY && carry != 0) { o.newc = carry; carry = 0; }
  } else if (action == A_TOGGLE) {                           // :576-578
    if (t4 == T4_DOOR_LOCKED) {                              // Door.toggle, world_object.py:184-194
      if ((carry & 15u) == T_KEY && ((carry >> 4) & 7u) == col) o.newc = T_DOOR | (col << 4);
    } else if (t4 == T_DOOR) o.newc = T4_DOOR_CLOSED | (col << 4) | OPAQUE_BIT;
    else if (t4 == T4_DOOR_CLOSED) o.newc = T_DOOR | (col << 4);
    else if (t4 == T_BOX) o.newc = CODE_EMPTY;               // Box.toggle, contains == None (:290-293)
  } else if (action != A_DONE) {
    o.bad_action = 1;
  }
  return o;
}

}  // namespace mg
