"""oracle/ -- TEST INFRASTRUCTURE ONLY.  CPU (numpy) restatement of the reference's hybrid-MPC hot path.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` (as the checker) and ``bench.py``'s
``cpu_baseline`` leg / ``--impl reference`` arm.  Nothing under ``pyhybridcontrol_b200/`` imports it; the
product path fails loudly when the CUDA library is missing.

What it restates (all citations relative to /root/reference):

* ``oracle.mld``       -- MLD dimension rules + default blocks   (models/mld_model.py:149-168, 515-520, 910-928)
* ``oracle.condense``  -- horizon condensing Phi/Gamma/L/H        (controllers/components/mld_evolution_matrices.py:237-332, 467-527)
* ``oracle.assemble``  -- variable layout, cost atoms, constraints (controllers/components/variables.py:189-243,
                          objective_atoms.py:308-363,453-496, controllers/controller_base.py:440-452,467-472)
* ``oracle.lsim``      -- one-step MLD simulation + DEWH sim model (models/mld_model.py:647-699,
                          examples/.../micro_grid_models.py:27-100, micro_grid_agents.py:389-408)
* ``oracle.callable``  -- symbolic / callable model evaluation: sympy.lambdify of every matrix, called per agent
                          (utils/matrix_utils.py:339-343, 372-380, 441-470; models/mld_model.py:791-793, 1128-1149)
* ``oracle.mini_cvxpy`` / ``oracle.ref_shim`` -- not restatements: the stand-in for cvxpy's modelling layer and
                          the compat shim that let the UNMODIFIED reference run here to produce the golden vectors
* ``oracle.solve``     -- the MI(Q)P solve the reference hands to cvxpy -> Gurobi/CPLEX
                          (controllers/controller_base.py:509-512).  Those solvers are third-party,
                          unpinned and not installed; the offline backend is HiGHS 1.12.0 as vendored by
                          scipy 1.18.1 (``scipy.optimize.milp``), plus exhaustive enumeration for <= ~18
                          binaries and a Python B&B over HiGHS QPs for quadratic costs.

Pinning status.  The reference has NO tests, golden vectors or fixtures for this path,
so per the task rules:

* condensing + lsim_k: **pinned** against outputs of the unmodified reference itself, run in the build
  container under ``oracle/ref_shim.py``; vectors committed in ``tests/golden/`` together with the
  generating script ``tests/golden/make_golden.py``.
* symbolic / callable models: **pinned** against the unmodified reference's CallableMatrix / MldSystemModel /
  DewhModel / GridModel / PvModel / ResDemandModel (``ref_shim.load_symbolic``); vectors ``tests/golden/callable_*.npz``,
  generating script ``tests/golden/make_golden_callable.py``.
* problem assembly, ``solve`` / ``feedback`` / ``sim_step_k`` flows: **pinned** against the unmodified reference's
  own code (EvoVariables, ObjectiveAtoms, gen_evo_constraints, MpcController.build / solve / feedback, sim_step_k with
  its auxiliary feasibility problem), run in the build container under ``ref_shim.load_controllers``: cvxpy's
  MODELLING layer is replaced by ``oracle/mini_cvxpy.py`` (cvxpy is third-party, not installed, unpinned by the
  reference -- API usage implies 1.0.x; the stand-in restates its documented semantics: column-major reshape, ``*``
  rules, atoms) and its MILP backend by HiGHS.  Vectors ``tests/golden/assembly_*.npz``, generating script
  ``tests/golden/make_golden_assembly.py``, checker ``tests/test_oracle_assembly_pinned.py``.
* the centralised micro-grid problem (``oracle.coupled``): **pinned** against the reference's own
  GridAgentMpc.build_grid / solve_grid_mpc / sim_step_k loop (``tests/golden/microgrid_loop.npz``,
  ``tests/golden/make_golden_microgrid.py``, ``tests/test_oracle_microgrid_pinned.py``).
* the mixed-integer SOLVER itself (Gurobi / CPLEX): **unpinned** -- not installable; HiGHS 1.12.0 (scipy) stands in,
  cross-checked against exhaustive enumeration.
* input side (profile windows, scenario draws, prices, tariff) and result frame: **pinned**
  (``tests/golden/profiles_inputs.npz``, ``simlog_campaign.npz``).
"""
