"""ORACLE — CPU restatement of the reference algorithm for the frustum hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``frustum_convnet_b200/`` may import this
package; it is used by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs as the checker / CPU baseline.
"""
