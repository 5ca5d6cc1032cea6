"""The descriptor channel of the peer-slab handshake (csrc/peer_ipc.cpp): two processes pass 40 file
descriptors over abstract Unix datagram sockets, the sender starting before the receiver is bound."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fd_channel_roundtrip(tmp_path):
    exe = str(tmp_path / "fd_channel_test")
    pkg = os.path.join(ROOT, "hiqsimulator_b200")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "fd_channel_test.cpp"),
                    "-o", exe, "-L" + pkg, "-lhiq_b200", "-Wl,-rpath," + pkg], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "FD_CHANNEL_OK" in res.stdout, (res.returncode, res.stdout, res.stderr)
