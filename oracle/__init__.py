"""CPU oracle (test infrastructure).  See yolo_nano_oracle.py."""
