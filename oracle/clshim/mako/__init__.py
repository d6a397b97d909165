"""ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for the one Mako feature the reference uses (see template.py)."""
