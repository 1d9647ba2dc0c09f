class HiddenStateNotInitialized(Exception):
    """Raised by BasicMAC.forward before init_hidden (reference: exceptions/mac_exceptions.py:1-3)."""

    def __init__(self):
        super().__init__("Please run init_hidden() to initialize the hidden state before running forward pass.)")
