"""Stand-in for matplotlib: pyplot calls of search_organisms.py:756-761 become no-ops."""
