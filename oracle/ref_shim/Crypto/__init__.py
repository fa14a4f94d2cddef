"""Stand-in for pycryptodome (pinned 3.9.9 in the reference's requirements.txt:144), which is
not installed in this image.  TEST INFRASTRUCTURE ONLY: it lets `tests/golden/make_golden.py`
import the *unmodified* reference modules from /root/reference inside the build container.
Nothing in the product package imports this."""
