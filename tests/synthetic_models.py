"""Synthetic MJCF models that exercise the kernel variants Ant and Humanoid do not
reach: a free torso with N two-hinge legs ending in a foot sphere."""
import math


def centipede_xml(n_legs: int, iterations: int = 6) -> str:
  legs = []
  for k in range(n_legs):
    ang = 2 * math.pi * k / n_legs
    dx, dy = 0.25 * math.cos(ang), 0.25 * math.sin(ang)
    ax, ay = -math.sin(ang), math.cos(ang)
    legs.append(f'''
      <body name="hip_{k}" pos="{dx:.4f} {dy:.4f} 0">
        <joint name="hz_{k}" axis="0 0 1" pos="0 0 0" range="-40 40" type="hinge"/>
        <geom fromto="0 0 0 {dx:.4f} {dy:.4f} 0" size="0.05" type="capsule"/>
        <body name="shin_{k}" pos="{dx:.4f} {dy:.4f} 0">
          <joint name="kn_{k}" axis="{ax:.4f} {ay:.4f} 0" pos="0 0 0" range="20 80" type="hinge"/>
          <geom fromto="0 0 0 {1.2 * dx:.4f} {1.2 * dy:.4f} 0" size="0.05" type="capsule"/>
          <geom name="foot_{k}" contype="1" pos="{1.2 * dx:.4f} {1.2 * dy:.4f} 0" size="0.06" type="sphere" mass="0"/>
        </body>
      </body>''')
  motors = ''.join(f'<motor joint="hz_{k}" gear="60" ctrlrange="-1 1"/><motor joint="kn_{k}" gear="60" ctrlrange="-1 1"/>'
                   for k in range(n_legs))
  return f'''
<mujoco model="centipede{n_legs}">
  <compiler angle="degree" inertiafromgeom="true"/>
  <option timestep="0.01" iterations="{iterations}"/>
  <custom><numeric data="15" name="solver_maxls"/><numeric data="8" name="matrix_inv_iterations"/></custom>
  <default>
    <joint armature="0.5" damping="1" limited="true"/>
    <geom contype="0" conaffinity="0" condim="3" density="20" friction="1 0.5 0.5"/>
  </default>
  <worldbody>
    <geom conaffinity="1" name="floor" pos="0 0 0" size="40 40 40" type="plane"/>
    <body name="torso" pos="0 0 0.5">
      <joint armature="0" damping="0" limited="false" name="root" type="free"/>
      <geom name="torso_geom" size="0.2" type="sphere"/>{''.join(legs)}
    </body>
  </worldbody>
  <actuator>{motors}</actuator>
</mujoco>'''


# Two kinematic trees in one System: a free ball and a free pendulum-carrying box, both
# touching the floor through spheres.  Exercises the per-tree root_com (dynamics.py:37-43),
# several free roots, and a contact-free tree next to a contacting one.
TWO_TREES_XML = '''
<mujoco model="two_trees">
  <compiler angle="radian" inertiafromgeom="true"/>
  <option timestep="0.005" iterations="6"/>
  <custom><numeric data="10" name="solver_maxls"/><numeric data="6" name="matrix_inv_iterations"/></custom>
  <default>
    <joint armature="0.2" damping="0.5" limited="true"/>
    <geom contype="0" conaffinity="0" condim="3" density="50" friction="0.8 0.1 0.1"/>
  </default>
  <worldbody>
    <geom conaffinity="1" name="floor" pos="0 0 0" size="40 40 40" type="plane"/>
    <body name="ball" pos="0 0 0.3">
      <joint armature="0" damping="0" limited="false" name="ball_root" type="free"/>
      <geom name="ball_geom" contype="1" size="0.25" type="sphere"/>
    </body>
    <body name="cart" pos="1.5 0 0.4">
      <joint armature="0" damping="0" limited="false" name="cart_root" type="free"/>
      <geom name="cart_geom" contype="1" size="0.3" type="sphere"/>
      <body name="arm" pos="0 0 0.3">
        <joint name="arm_hinge" axis="0 1 0" pos="0 0 0" range="-1.2 1.2" type="hinge"/>
        <geom fromto="0 0 0 0 0 0.5" size="0.05" type="capsule"/>
        <body name="tip" pos="0 0 0.5">
          <joint name="tip_hinge" axis="1 0 0" pos="0 0 0" range="-0.8 0.8" type="hinge"/>
          <geom name="tip_geom" size="0.1" type="sphere"/>
        </body>
      </body>
    </body>
  </worldbody>
  <actuator><motor joint="arm_hinge" gear="20" ctrlrange="-1 1"/><motor joint="tip_hinge" gear="10" ctrlrange="-1 1"/></actuator>
</mujoco>'''
